"""Golden vectors for the training-side use of the path, from the REFERENCE ITSELF.

    python tests/golden/make_golden_train.py        (build container only)

Runs the reference's `q_pred`, `q_pred_one_timestep`, `q_sample` and `_train_loss` (with its `sample_time`
replaced by fixed timesteps and `torch.rand_like` by an injected uniform tensor) and a backward pass of the loss
that `forward()` forms (`loss.sum() / (B*N)`, :554) with respect to the denoiser logits.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import d3pm_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
T = 100


def tm(x):
    return np.ascontiguousarray(x.detach().permute(0, 2, 1).numpy())


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    # ---- forward process operators on a non-one-hot and a one-hot input
    B, N, K = 3, 8, 64
    sched = O.make_schedule(T, K)
    lc, lu, x_t, _, u = O.synth_inputs(B, N, K, torch.tensor([0, 42, 99]), sched, seed=300)
    ref, model = R.make_reference_model(K, T, N, 2.0, lc, lu)
    t = torch.tensor([0, 42, 99])
    g = torch.Generator().manual_seed(301)
    x0 = torch.randint(0, K, (B, N), generator=g)
    soft = O.predict_start_from_logits(lc.permute(0, 2, 1))
    hot = ref.index_to_log_onehot(x0, K + 1)
    with torch.no_grad():
        outs = dict(qpred_soft=model.q_pred(soft, t), qpred_hot=model.q_pred(hot, t), qpred_hot_tm1=model.q_pred(hot, t - 1),
                    qone_hot=model.q_pred_one_timestep(ref.index_to_log_onehot(x_t, K + 1), t),
                    qone_soft=model.q_pred_one_timestep(soft, t))
        with R.injected_uniform(u):
            xs = model.q_sample(hot, t).argmax(1)
    np.savez_compressed(os.path.join(OUT, "forward_process_k64.npz"), soft=tm(soft), x0=x0.numpy(), x_t=x_t.numpy(),
                        t=t.numpy(), uniform=tm(u), q_sample=xs.numpy(), T=np.int32(T), K=np.int32(K),
                        **{k: tm(v) for k, v in outs.items()})
    print("forward process:", {k: tuple(v.shape) for k, v in outs.items()})

    # ---- the training loss, forward and backward, two configurations
    for name, K, B, N, tvals, aux, adaptive, scale, seed in [
        ("k64_aux", 64, 4, 8, [0, 3, 57, 99], 5.0e-4, True, 1.0, 400),
        ("k64_noaux_stress", 64, 3, 8, [1, 50, 98], 0.0, False, 8.0, 410),
        ("k2048", 2048, 2, 3, [20, 0], 5.0e-4, True, 1.0, 420),
    ]:
        sched = O.make_schedule(T, K)
        lc, lu, _, _, u = O.synth_inputs(B, N, K, 50, sched, seed=seed, scale=scale, spikes=scale > 1)
        logits = lc.clone().requires_grad_(True)
        ref, model = R.make_reference_model(K, T, N, 2.0, logits, logits)
        model.auxiliary_loss_weight, model.adaptive_auxiliary_loss = aux, adaptive
        g = torch.Generator().manual_seed(seed + 1)
        x0 = torch.randint(0, K, (B, N), generator=g)
        t = torch.tensor(tvals)
        pt = torch.ones(B) / T
        model.sample_time = lambda b, device, method="uniform": (t, pt)
        with R.injected_uniform(u):
            out = model({"content_token": x0, "condition_embed_token": torch.ones(B, 1, 512)}, return_loss=True)
        out["loss"].backward()
        with torch.no_grad(), R.injected_uniform(u):
            log_model_prob, vb_loss, x0_recon = model._train_loss(x0, torch.ones(B, 1, 512))
        xt = O.q_sample(sched, O.index_to_log_onehot(x0, K + 1), t, u).argmax(1)
        np.savez_compressed(
            os.path.join(OUT, f"train_loss_{name}.npz"), logits=lc.numpy(), x0=x0.numpy(), t=t.numpy(), pt=pt.numpy(),
            uniform=tm(u), x_t=xt.numpy(), log_model_prob=tm(log_model_prob), vb_loss=vb_loss.numpy(),
            x0_recon=x0_recon.numpy(), loss=out["loss"].detach().numpy(), grad_logits=logits.grad.numpy(),
            probs=tm(out["logits"]), T=np.int32(T), K=np.int32(K), aux=np.float32(aux), adaptive=np.bool_(adaptive))
        print(name, "loss", float(out["loss"]), "grad max", float(logits.grad.abs().max()))


if __name__ == "__main__":
    main()

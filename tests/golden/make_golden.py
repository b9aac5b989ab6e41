"""Generate the committed golden vectors from the REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src/models/motionencoder/diffusion_transformer.py by path
(oracle/ref_loader.py), runs the reference's own `p_pred`, `p_sample`, `q_posterior`,
`predict_start`, `log_sample_categorical` and `sample` on seeded synthetic inputs
(oracle.d3pm_oracle.synth_inputs, SURVEY.md §8 d) with the uniform noise injected through
`torch.rand_like`, and writes `tests/golden/*.npz`.  The reference holds no tests or
vectors for this path, so these files are what pins the oracle and the CUDA path.

Everything is stored token-major ([B,N,K+1]); the reference's logical layout is
[B,K+1,N] and is permuted on the way out.
"""
from __future__ import annotations

import hashlib
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

# name, B, N, K, t, logit scale, spikes, guidance scale (None = guidance off), seed
SMALL_CASES = [
    ("k64_t50", 2, 8, 64, 50, 1.0, False, 2.0, 0),
    ("k64_t0", 2, 8, 64, 0, 1.0, False, 2.0, 10),
    ("k64_t99_stress", 2, 8, 64, 99, 8.0, True, 2.0, 20),
    ("k64_t1", 2, 8, 64, 1, 1.0, False, 2.0, 30),
    ("k64_pert_s5_stress", 3, 8, 64, [0, 1, 57], 8.0, True, 5.0, 40),
    ("k64_t50_noguid", 2, 8, 64, 50, 1.0, False, None, 50),
    ("k64_t0_noguid_stress", 2, 8, 64, 0, 8.0, True, None, 60),
    ("k2048_t25", 1, 3, 2048, 25, 1.0, False, 2.0, 70),
    ("k4096_t50", 1, 4, 4096, 50, 1.0, False, 2.0, 80),
    ("k4096_t0_stress", 1, 4, 4096, 0, 8.0, True, 2.0, 90),
]


def sha(*tensors) -> str:
    h = hashlib.sha256()
    for x in tensors:
        h.update(np.ascontiguousarray(x.numpy()).tobytes())
    return h.hexdigest()


def tm(x: torch.Tensor) -> np.ndarray:  # [B,C,N] -> token-major numpy
    return np.ascontiguousarray(x.permute(0, 2, 1).numpy())


def run_reference_step(B, N, K, tval, scale, spikes, s, seed):
    sched = O.make_schedule(T, K)
    t_in = torch.tensor(tval) if isinstance(tval, list) else tval
    lc, lu, x_t, t, u = O.synth_inputs(B, N, K, t_in, sched, seed=seed, scale=scale, spikes=spikes)
    ref, model = R.make_reference_model(K, T, N, 2.0 if s is None else s, lc, lu)
    for name in O.SCHEDULE_NAMES:  # the oracle's schedule must be the reference's, bit for bit
        assert torch.equal(getattr(model, name), sched[name]), name
    log_x_t = ref.index_to_log_onehot(x_t, K + 1)
    cond, cf = torch.ones(B, 1, 512), torch.zeros(B, 1, 512)
    with torch.no_grad():
        if s is None:  # guidance off: predict_start -> q_posterior -> sampler
            recon = model.predict_start(log_x_t, cond, t)
            post = model.q_posterior(recon, log_x_t, t)
            with R.injected_uniform(u):
                out = model.log_sample_categorical(post)
        else:
            post, recon = model.p_pred(log_x_t, cond, cf, t)
            with R.injected_uniform(u):
                out, sampled = model.p_sample(log_x_t, cond, cf, t, [0] * B, 10)
            assert sampled == [1024] * B
    tok = out.argmax(1)
    ties = O.near_ties(post, u)
    return dict(lc=lc, lu=lu, x_t=x_t, t=t, u=u, post=post, recon=recon, tok=tok, ties=ties), (ref, model, sched)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for name, B, N, K, tval, scale, spikes, s, seed in SMALL_CASES:
        r, _ = run_reference_step(B, N, K, tval, scale, spikes, s, seed)
        np.savez_compressed(
            os.path.join(OUT, f"step_{name}.npz"),
            logits_c=r["lc"].numpy(), logits_u=r["lu"].numpy(), x_t=r["x_t"].numpy(), t=r["t"].numpy(),
            uniform=tm(r["u"]), post=tm(r["post"]), recon=tm(r["recon"]), x_prev=r["tok"].numpy(),
            near_tie=r["ties"].numpy(), guidance_scale=np.float32(-1.0 if s is None else s),
            T=np.int32(T), K=np.int32(K))
        print(name, "clamped", float((r["post"] <= -70).float().mean()), "ties", int(r["ties"].sum()))

    # standalone q_posterior on a one-hot x_0 (the training-side call, :420) with per-sample t
    B, N, K = 3, 8, 64
    sched = O.make_schedule(T, K)
    lc, lu, x_t, t, u = O.synth_inputs(B, N, K, torch.tensor([3, 42, 99]), sched, seed=100)
    ref, model = R.make_reference_model(K, T, N, 2.0, lc, lu)
    g = torch.Generator().manual_seed(104)
    x0 = torch.randint(0, K, (B, N), generator=g)
    with torch.no_grad():
        post = model.q_posterior(ref.index_to_log_onehot(x0, K + 1), ref.index_to_log_onehot(x_t, K + 1), t)
    np.savez_compressed(os.path.join(OUT, "qpost_onehot_k64.npz"), x0=x0.numpy(), x_t=x_t.numpy(), t=t.numpy(),
                        post=tm(post), T=np.int32(T), K=np.int32(K))

    # a whole sample() loop (T=10 is one of the step counts update_n_sample knows, :166-179)
    B, N, K, Ts = 2, 8, 64, 10
    sched = O.make_schedule(Ts, K)
    lc, lu, _, _, _ = O.synth_inputs(B, N, K, 5, sched, seed=200)
    ref, model = R.make_reference_model(K, Ts, N, 2.0, lc, lu)
    g = torch.Generator().manual_seed(203)
    us = [torch.rand(B, K + 1, N, generator=g) for _ in range(Ts)]
    it = iter(us)
    real = torch.rand_like
    trace = []
    orig_p_sample = model.p_sample

    def traced(*a, **k):
        out, sampled = orig_p_sample(*a, **k)
        trace.append(out.argmax(1).clone())
        return out, sampled

    model.p_sample = traced
    torch.rand_like = lambda x, *a, **k: next(it)
    try:
        res = model.sample(["a"] * B, None, torch.ones(B, 1, 512), torch.zeros(B, 1, 512),
                           content_token=None, filter_ratio=0)
    finally:
        torch.rand_like = real
    np.savez_compressed(os.path.join(OUT, "sample_loop_k64_T10.npz"), logits_c=lc.numpy(), logits_u=lu.numpy(),
                        uniforms=np.stack([tm(x) for x in us]), trace=torch.stack(trace).numpy(),
                        content_token=res["content_token"].numpy(), T=np.int32(Ts), K=np.int32(K),
                        guidance_scale=np.float32(2.0))
    print("sample loop", res["content_token"].shape, "denoiser calls", model.transformer.calls)

    # config 1 at full size (B=1, 16x8x8 grid, K=4096): inputs are regenerated from seeds by the
    # tests, so only digests, tokens and a posterior digest/sample are stored
    B, N, K, seed = 1, 1024, 4096, 0
    r, _ = run_reference_step(B, N, K, 50, 1.0, False, 2.0, seed)
    post_tm = r["post"].permute(0, 2, 1).contiguous()
    rows = np.arange(0, N, 64)
    np.savez_compressed(
        os.path.join(OUT, "step_config1_digest.npz"),
        inputs_sha256=np.array(sha(r["lc"], r["lu"], r["x_t"], r["t"], r["u"])),
        x_prev=r["tok"].numpy(), near_tie=r["ties"].numpy(), sample_rows=rows,
        post_rows=post_tm[0, rows].numpy(),
        post_row_lse=torch.logsumexp(post_tm.double(), -1).numpy(),
        post_row_sum=post_tm.double().sum(-1).numpy(),
        seed=np.int32(seed), T=np.int32(T), K=np.int32(K), guidance_scale=np.float32(2.0))
    print("config1 ties", int(r["ties"].sum()), "masked", int((r["x_t"] == K).sum()))


if __name__ == "__main__":
    main()

"""Golden vectors for rows where the denoiser is SURE of x_t (a logit gap > 16 on the current token) at small t.

    python tests/golden/make_golden_confident.py        (build container: needs /root/reference)

There p(x0 = x_t | x_t) is within 1e-7 of 1 while 1 / b_t is ~1e8, so a closed form that writes the mass of the other
classes as 1 - p_j loses it to cancellation; the reference works class by class in the log domain (:251-283).  The
reference's own `p_pred` (guidance 2) and `predict_start` + `q_posterior` (guidance off) on such rows, t in {0, 1, 3, 5}.
"""
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
    return np.ascontiguousarray(x.permute(0, 2, 1).numpy())


def make(name, K, s, seed):
    B, N = 4, 8
    sched = O.make_schedule(T, K)
    t = torch.tensor([0, 1, 3, 5])
    lc, lu, x_t, _, u = O.synth_inputs(B, N, K, t, sched, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x_t = torch.randint(0, K, (B, N), generator=g)   # every token unmasked ...
    x_t[:, 0] = K                                     # ... but one [MASK] per video
    gap = 16.5 + 6.0 * torch.rand(B, N, generator=g)
    for b in range(B):
        for n in range(1, N):
            lc[b, n, x_t[b, n]] = lc[b, n].max() + gap[b, n]
            lu[b, n, x_t[b, n]] = lu[b, n].max() + gap[b, n] * 0.9
    ref, model = R.make_reference_model(K, T, N, 2.0 if s is None else s, lc, lu)
    log_x_t = ref.index_to_log_onehot(x_t, K + 1)
    cond, cf = torch.ones(B, 1, 512), torch.zeros(B, 1, 512)
    with torch.no_grad():
        if s is None:
            recon = model.predict_start(log_x_t, cond, t)
            post = model.q_posterior(recon, log_x_t, t)
            with R.injected_uniform(u):
                out = model.log_sample_categorical(post)
        else:
            post, recon = model.p_pred(log_x_t, cond, cf, t)
            with R.injected_uniform(u):
                out, _ = model.p_sample(log_x_t, cond, cf, t, [0] * B, 10)
    tok = out.argmax(1)
    ties = O.near_ties(post, u)
    # the same keys as tests/golden/make_golden.py writes: the file joins the parametrised step-fixture tests
    np.savez_compressed(os.path.join(OUT, f"step_{name}.npz"), logits_c=lc.numpy(), logits_u=lu.numpy(), x_t=x_t.numpy(),
                        t=t.numpy(), uniform=tm(u), post=tm(post), recon=tm(recon), x_prev=tok.numpy(), near_tie=ties.numpy(),
                        guidance_scale=np.float32(-1.0 if s is None else s), T=np.int32(T), K=np.int32(K))
    off = post.clone()
    off.scatter_(1, x_t.clamp(max=K).unsqueeze(1), 0.0)
    print(name, "p_j max", float(recon.exp().max()), "off-class posterior range", float(off[:, :K].min()), float(off[:, :K].max()),
          "ties", int(ties.sum()))


if __name__ == "__main__":
    make("k64_confident", 64, 2.0, 300)
    make("k64_confident_noguid", 64, None, 310)

"""Golden vectors for the purity-prior branch of the reference's `p_sample` (prior_rule 1 / 2, :304-352) — build
container only.

    python tests/golden/make_golden_purity.py

`prior_rule` is hard-coded to 0 in the reference's constructor (:157), so the branch is reached by setting the
attribute on the imported reference model.  The two noise sources are injected: `torch.rand_like` (Gumbel draw,
:355) and the Exp(1) tensor inside `torch.multinomial` (:340).  The latter is injected by replacing
`torch.multinomial` with `topk(weights / q, n)`, after checking here that this IS what the real op computes.
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

# name, B, N, K, t, prior_rule, prior_weight, to_sample, sampled-before, seed
CASES = [
    ("rule2_w0", 3, 64, 64, 60, 2, 0.0, 11, [0, 0, 0], 300),
    ("rule2_w1p5", 3, 64, 64, 40, 2, 1.5, 10, [0, 3, 10], 310),
    ("rule1", 2, 64, 64, 70, 1, 0.0, 6, [0, 5], 320),
    ("rule2_k4096", 2, 16, 4096, 60, 2, 0.0, 5, [0, 1], 330),
    ("rule2_k4096_w2", 1, 16, 4096, 70, 2, 2.0, 5, [0], 340),
]


def check_multinomial_equivalence():
    w = torch.rand(257)
    w[::3] = 0
    for n in (1, 4, 33):
        g = torch.Generator().manual_seed(5)
        real = torch.multinomial(w, n, generator=g)
        g = torch.Generator().manual_seed(5)
        q = torch.empty_like(w).exponential_(1, generator=g)
        mine = (w / q).argmax(-1, keepdim=True) if n == 1 else torch.topk(w / q, n).indices
        assert torch.equal(real, mine), "torch.multinomial is no longer topk(weights / Exp(1))"


def main():
    check_multinomial_equivalence()
    for name, B, N, K, tval, rule, weight, to_sample, sampled0, seed in CASES:
        sched = O.make_schedule(T, K)
        lc, lu, x_t, t, u = O.synth_inputs(B, N, K, tval, sched, seed=seed)
        g = torch.Generator().manual_seed(seed + 7)
        expo = torch.empty(B, N).exponential_(1, generator=g)
        ref, model = R.make_reference_model(K, T, N, 2.0, lc, lu)
        model.prior_rule, model.prior_weight, model.prior_ps = rule, weight, 1024
        # keep the draw well defined: each video must hold at least as many [MASK] tokens as it is asked to reveal
        # (otherwise topk falls into the zero-weight ties, whose order is unspecified)
        assert all(int((x_t[i] == K).sum()) >= to_sample - sampled0[i] for i in range(B)), name
        log_x_t = ref.index_to_log_onehot(x_t, K + 1)
        cond, cf = torch.ones(B, 1, 512), torch.zeros(B, 1, 512)
        real_multinomial = torch.multinomial

        def fake_multinomial(w, n, *a, **k):
            i = w.storage_offset() // N  # `_score[i]` is a row view of the [B, N] score matrix
            return torch.topk(w / expo[i], n).indices

        torch.multinomial = fake_multinomial
        try:
            with torch.no_grad(), R.injected_uniform(u):
                out, sampled = model.p_sample(log_x_t, cond, cf, t, list(sampled0), to_sample)
        finally:
            torch.multinomial = real_multinomial
        tok = out.argmax(1)
        tok_o, sampled_o = O.p_sample_purity_step(sched, lc.permute(0, 2, 1), lu.permute(0, 2, 1), log_x_t, t, 2.0, u, expo,
                                                  sampled0, to_sample, prior_rule=rule, prior_weight=weight)
        assert torch.equal(tok, tok_o) and [int(s) for s in sampled] == sampled_o, name
        np.savez_compressed(
            os.path.join(OUT, f"purity_{name}.npz"), logits_c=lc.numpy(), logits_u=lu.numpy(), x_t=x_t.numpy(),
            t=t.numpy(), uniform=np.ascontiguousarray(u.permute(0, 2, 1).numpy()), expo=expo.numpy(),
            x_prev=tok.numpy(), sampled_in=np.array(sampled0), sampled_out=np.array([int(s) for s in sampled]),
            to_sample=np.int32(to_sample), prior_rule=np.int32(rule), prior_weight=np.float32(weight),
            guidance_scale=np.float32(2.0), T=np.int32(T), K=np.int32(K))
        print(name, "changed", int((tok != x_t).sum()), "sampled", sampled)


if __name__ == "__main__":
    main()

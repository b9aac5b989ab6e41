"""Shared helpers for the tests (the oracle is imported here as the CHECKER only)."""
import glob
import os

import numpy as np
import torch

from oracle import d3pm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
POST_TOL = 1e-4       # north_star: posterior log-probs within 1e-4 absolute in fp32
NEAR_TIE_GAP = 2e-4   # tokens must match except where the reference's own top-2 gap is below this


def step_fixtures():
    return sorted(glob.glob(os.path.join(GOLDEN, "step_k*.npz")))


def load(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def fixture_guidance(fx):
    s = float(fx["guidance_scale"])
    return None if s < 0 else s


def oracle_step_from_fixture(fx):
    """Run the op-faithful oracle on a fixture's inputs -> (tokens [B,N], post [B,N,K+1], recon [B,N,K+1])."""
    T, K = int(fx["T"]), int(fx["K"])
    sched = O.make_schedule(T, K)
    lc, lu = torch.from_numpy(fx["logits_c"]), torch.from_numpy(fx["logits_u"])
    x_t, t = torch.from_numpy(fx["x_t"]), torch.from_numpy(fx["t"])
    u = torch.from_numpy(fx["uniform"]).permute(0, 2, 1)
    s = fixture_guidance(fx)
    out, post, recon = O.p_sample_step(sched, lc.permute(0, 2, 1), None if s is None else lu.permute(0, 2, 1),
                                       O.index_to_log_onehot(x_t, K + 1), t, 0.0 if s is None else s, u)
    return out.argmax(1), post.permute(0, 2, 1), recon.permute(0, 2, 1)


def assert_tokens_match(got, want, near_tie, what=""):
    """Bit-exact except at (logged) near-ties."""
    got, want, near_tie = np.asarray(got), np.asarray(want), np.asarray(near_tie).astype(bool)
    diff = got != want
    if diff.any():
        print(f"[near-tie log] {what}: {int(diff.sum())} token(s) differ, {int((diff & near_tie).sum())} of them at near-ties; "
              f"{int(near_tie.sum())} near-tie position(s) in total")
    assert not (diff & ~near_tie).any(), f"{what}: {int((diff & ~near_tie).sum())} token mismatches away from near-ties"

"""Purity-prior branch of p_sample: the oracle against the reference's golden outputs (CPU)."""
import glob

import numpy as np
import pytest
import torch

from oracle import d3pm_oracle as O
from tests import helpers as H

FIXTURES = sorted(glob.glob(f"{H.GOLDEN}/purity_*.npz"))


def test_fixtures_present():
    assert len(FIXTURES) >= 5


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: p.split("purity_")[-1][:-4])
def test_oracle_reproduces_reference(path):
    fx = H.load(path)
    T, K = int(fx["T"]), int(fx["K"])
    sched = O.make_schedule(T, K)
    lc, lu = torch.from_numpy(fx["logits_c"]).permute(0, 2, 1), torch.from_numpy(fx["logits_u"]).permute(0, 2, 1)
    x_t, t = torch.from_numpy(fx["x_t"]), torch.from_numpy(fx["t"])
    u = torch.from_numpy(fx["uniform"]).permute(0, 2, 1)
    tok, sampled = O.p_sample_purity_step(sched, lc, lu, O.index_to_log_onehot(x_t, K + 1), t, float(fx["guidance_scale"]), u,
                                          torch.from_numpy(fx["expo"]), fx["sampled_in"].tolist(), int(fx["to_sample"]),
                                          prior_rule=int(fx["prior_rule"]), prior_weight=float(fx["prior_weight"]))
    assert np.array_equal(tok.numpy(), fx["x_prev"])
    assert sampled == fx["sampled_out"].tolist()
    # what the branch guarantees: only [MASK] positions change, and exactly the requested number per video
    changed = tok != x_t
    assert bool((x_t[changed] == K).all())
    want = [min(int(fx["to_sample"]) - s, 1024) for s in fx["sampled_in"].tolist()]
    assert changed.sum(1).tolist() == [max(w, 0) for w in want]


def test_multinomial_is_the_exponential_race():
    """The noise-injection point: torch.multinomial(w, n) without replacement == topk(w / Exp(1))."""
    w = torch.rand(300)
    w[::4] = 0
    for n in (1, 7, 50):
        g = torch.Generator().manual_seed(11)
        real = torch.multinomial(w, n, generator=g)
        g = torch.Generator().manual_seed(11)
        q = torch.empty_like(w).exponential_(1, generator=g)
        mine = (w / q).argmax(-1, keepdim=True) if n == 1 else O.multinomial_without_replacement(w, n, q)
        assert torch.equal(real, mine)

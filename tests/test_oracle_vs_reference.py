"""Oracle against the live reference (build container only; skipped where /root/reference is absent)."""
import numpy as np
import pytest
import torch

from oracle import d3pm_oracle as O
from oracle import ref_loader as R

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="reference tree not present on this machine")


@pytest.mark.parametrize("seed,tval,scale,spikes,s", [(1000, 73, 1.0, False, 2.0), (1010, 0, 4.0, True, 3.0),
                                                      (1020, [5, 99], 1.0, False, 2.0)])
def test_oracle_equals_reference_on_fresh_seeds(seed, tval, scale, spikes, s):
    B, N, K, T = 2, 6, 128, 100
    sched = O.make_schedule(T, K)
    t_in = torch.tensor(tval) if isinstance(tval, list) else tval
    lc, lu, x_t, t, u = O.synth_inputs(B, N, K, t_in, sched, seed=seed, scale=scale, spikes=spikes)
    ref, model = R.make_reference_model(K, T, N, s, lc, lu)
    for name in O.SCHEDULE_NAMES:
        assert torch.equal(getattr(model, name), sched[name])
    log_x_t = ref.index_to_log_onehot(x_t, K + 1)
    cond, cf = torch.ones(B, 1, 512), torch.zeros(B, 1, 512)
    with torch.no_grad():
        post_ref, recon_ref = model.p_pred(log_x_t, cond, cf, t)
        with R.injected_uniform(u):
            out_ref, _ = model.p_sample(log_x_t, cond, cf, t, [0] * B, 10)
    out, post, recon = O.p_sample_step(sched, lc.permute(0, 2, 1), lu.permute(0, 2, 1),
                                       O.index_to_log_onehot(x_t, K + 1), t, s, u)
    assert torch.equal(post, post_ref) and torch.equal(recon, recon_ref)
    assert torch.equal(out.argmax(1), out_ref.argmax(1))

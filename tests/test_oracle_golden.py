"""The oracle against the golden vectors generated from the reference itself (CPU, no GPU)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import d3pm_oracle as O
from tests import helpers as H


@pytest.mark.parametrize("path", H.step_fixtures(), ids=lambda p: p.split("step_")[-1][:-4])
def test_op_faithful_oracle_reproduces_reference(path):
    fx = H.load(path)
    tok, post, recon = H.oracle_step_from_fixture(fx)
    assert np.array_equal(tok.numpy(), fx["x_prev"])                    # sampled indices: bit-exact
    assert np.abs(post.numpy() - fx["post"]).max() <= 1e-6              # same op sequence -> same floats
    assert np.abs(recon.numpy() - fx["recon"]).max() <= 1e-6


@pytest.mark.parametrize("path", H.step_fixtures(), ids=lambda p: p.split("step_")[-1][:-4])
def test_closed_form_matches_reference(path):
    """The float64 per-token algebra the CUDA kernel implements, against the reference's outputs."""
    fx = H.load(path)
    if int(fx["K"]) > 2048:
        fx = {k: (v[:, :2] if k in ("logits_c", "logits_u", "x_t", "post", "recon") else v) for k, v in fx.items()}
    sched = O.make_schedule(int(fx["T"]), int(fx["K"]))
    s = H.fixture_guidance(fx)
    recon, post = O.closed_form_rows(sched, fx["logits_c"], None if s is None else fx["logits_u"], fx["x_t"], fx["t"],
                                     0.0 if s is None else s)
    assert np.abs(post - fx["post"]).max() <= H.POST_TOL
    assert np.abs(recon - fx["recon"]).max() <= H.POST_TOL


def test_qposterior_onehot_fixture():
    fx = H.load(f"{H.GOLDEN}/qpost_onehot_k64.npz")
    T, K = int(fx["T"]), int(fx["K"])
    sched = O.make_schedule(T, K)
    post = O.q_posterior(sched, O.index_to_log_onehot(torch.from_numpy(fx["x0"]), K + 1),
                         O.index_to_log_onehot(torch.from_numpy(fx["x_t"]), K + 1), torch.from_numpy(fx["t"]))
    assert np.abs(post.permute(0, 2, 1).numpy() - fx["post"]).max() <= 1e-6


def test_sample_loop_fixture():
    """The reference's whole sample() chain (T=10) replayed step by step through the oracle."""
    fx = H.load(f"{H.GOLDEN}/sample_loop_k64_T10.npz")
    T, K = int(fx["T"]), int(fx["K"])
    sched = O.make_schedule(T, K)
    lc, lu = torch.from_numpy(fx["logits_c"]).permute(0, 2, 1), torch.from_numpy(fx["logits_u"]).permute(0, 2, 1)
    B, _, N = lc.shape
    x = torch.full((B, N), K, dtype=torch.long)
    for i, step in enumerate(range(T - 1, -1, -1)):
        t = torch.full((B,), step, dtype=torch.long)
        u = torch.from_numpy(fx["uniforms"][i]).permute(0, 2, 1)
        out, _, _ = O.p_sample_step(sched, lc, lu, O.index_to_log_onehot(x, K + 1), t, float(fx["guidance_scale"]), u)
        x = out.argmax(1)
        assert np.array_equal(x.numpy(), fx["trace"][i])
    assert np.array_equal(x.numpy(), fx["content_token"])
    assert not (x == K).any()  # nothing is left masked after t = 0


def test_config1_digest():
    """Config 1 (B=1, 16x8x8 grid, K=4096) at full size: inputs regenerate bit-identically from the seeds
    and the oracle reproduces the reference's tokens and posterior."""
    fx = H.load(f"{H.GOLDEN}/step_config1_digest.npz")
    T, K = int(fx["T"]), int(fx["K"])
    sched = O.make_schedule(T, K)
    lc, lu, x_t, t, u = O.synth_inputs(1, 1024, K, 50, sched, seed=int(fx["seed"]))
    h = hashlib.sha256()
    for a in (lc, lu, x_t, t, u):
        h.update(np.ascontiguousarray(a.numpy()).tobytes())
    assert h.hexdigest() == str(fx["inputs_sha256"])
    out, post, _ = O.p_sample_step(sched, lc.permute(0, 2, 1), lu.permute(0, 2, 1), O.index_to_log_onehot(x_t, K + 1),
                                   t, float(fx["guidance_scale"]), u)
    assert np.array_equal(out.argmax(1).numpy(), fx["x_prev"])
    post_tm = post.permute(0, 2, 1)
    assert np.abs(post_tm[0, fx["sample_rows"]].numpy() - fx["post_rows"]).max() <= 1e-6
    assert np.abs(torch.logsumexp(post_tm.double(), -1).numpy() - fx["post_row_lse"]).max() <= 1e-6


def test_schedule_identity_slot_and_normalisation():
    sched = O.make_schedule(100, 4096)
    assert sched["log_cumprod_at"][-1] == 0 and torch.isneginf(sched["log_cumprod_bt"][-1])
    assert torch.isneginf(sched["log_cumprod_ct"][-1])
    tot = O.log_add_exp(sched["log_ct"].double(), sched["log_1_min_ct"].double())
    assert tot.abs().sum() < 1e-5  # the reference's own in-code assertion (:136)

"""Purity-prior sampling (p_sample with prior_rule 1 / 2) on the GPU against the reference's golden
outputs and the oracle.  Needs a B200."""
import glob
import types

import numpy as np
import pytest
import torch

import d3pm_b200
from d3pm_b200 import _lib, ops
from oracle import d3pm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FIXTURES = sorted(glob.glob(f"{H.GOLDEN}/purity_*.npz"))


class StubDenoiser(torch.nn.Module):
    def __init__(self, K, lc, lu):
        super().__init__()
        self.content_emb = types.SimpleNamespace(num_embed=K + 1)
        self.lc, self.lu, self.calls = lc, lu, 0

    def forward(self, x_t, cond, t):
        self.calls += 1
        return (self.lc if float(cond.flatten()[0]) > 0.5 else self.lu).permute(0, 2, 1)


def _model(fx):
    K, T = int(fx["K"]), int(fx["T"])
    lc, lu = torch.from_numpy(fx["logits_c"]).to(DEV), torch.from_numpy(fx["logits_u"]).to(DEV)
    m = d3pm_b200.FusedDiffusionTransformer(transformer=StubDenoiser(K, lc, lu), diffusion_step=T, alpha_init_type="alpha1",
                                            guidance_scale=float(fx["guidance_scale"]), content_seq_len=lc.shape[1]).to(DEV)
    B = lc.shape[0]
    return m, torch.ones(B, 1, 512, device=DEV), torch.zeros(B, 1, 512, device=DEV)


def _near_ties(fx):
    """Positions whose candidate draw is a near-tie in the reference's own arithmetic."""
    K, T = int(fx["K"]), int(fx["T"])
    lc, lu = torch.from_numpy(fx["logits_c"]).permute(0, 2, 1), torch.from_numpy(fx["logits_u"]).permute(0, 2, 1)
    recon = O.cf_predict_start_from_logits(lc, lu, float(fx["guidance_scale"]))
    rule, w = int(fx["prior_rule"]), float(fx["prior_weight"])
    prob = recon
    if rule != 1 and w > 0:
        score = torch.exp(recon).max(dim=1).values.clamp(0, 1)
        score = score / (score.max(dim=1, keepdim=True).values + 1e-10)
        prob = ((1 + score * w).unsqueeze(1) * recon).softmax(dim=1).log().clamp(-70, 0)
    return O.near_ties(prob, torch.from_numpy(fx["uniform"]).permute(0, 2, 1)).numpy()


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: p.split("purity_")[-1][:-4])
def test_p_sample_purity_against_golden(path):
    fx = H.load(path)
    K = int(fx["K"])
    m, cond, cf = _model(fx)
    m.prior_rule, m.prior_weight = int(fx["prior_rule"]), float(fx["prior_weight"])
    x_t, t = torch.from_numpy(fx["x_t"]).to(DEV), torch.from_numpy(fx["t"]).to(DEV)
    u = torch.from_numpy(fx["uniform"]).permute(0, 2, 1).contiguous()
    expo = torch.from_numpy(fx["expo"])
    m.inject_uniform = lambda shape, dev: u.to(dev)
    m.inject_exponential = lambda shape, dev: expo.to(dev)
    log_x = ops.as_logical(ops.tokens_to_log_onehot_rows(x_t, K + 1), K + 1)
    out, sampled = m.p_sample(log_x, cond, cf, t, fx["sampled_in"].tolist(), int(fx["to_sample"]))
    tok = out.argmax(1).cpu().numpy()
    # the revealed SET is exact; the revealed tokens match except at logged near-ties of the candidate draw
    assert np.array_equal(tok != fx["x_t"], fx["x_prev"] != fx["x_t"])
    H.assert_tokens_match(tok, fx["x_prev"], _near_ties(fx), path)
    assert sampled == fx["sampled_out"].tolist()
    m.check_status()


def test_purity_score_and_candidates_against_oracle():
    """The fused pass alone: purity = max_k p(x0 = k | x_t), candidates = Gumbel-max over log_x_recon (all K+1 classes)."""
    fx = H.load(f"{H.GOLDEN}/purity_rule2_w0.npz")
    K, T = int(fx["K"]), int(fx["T"])
    sched = O.make_schedule(T, K)
    lc, lu = torch.from_numpy(fx["logits_c"]), torch.from_numpy(fx["logits_u"])
    x_t, t = torch.from_numpy(fx["x_t"]), torch.from_numpy(fx["t"])
    table = ops.build_coef_table(O.pack_schedule(sched).to(DEV), T, K)
    recon = O.cf_predict_start_from_logits(lc.permute(0, 2, 1), lu.permute(0, 2, 1), 2.0)
    u = torch.from_numpy(fx["uniform"])
    urows = ops.alloc_rows(*x_t.shape, K + 1, DEV)
    urows[:, :, :K + 1] = u.to(DEV)
    out = ops.fused_step(lc.to(DEV), lu.to(DEV), x_t.to(DEV), t.to(DEV), table, guidance_scale=2.0,
                         sample_mode=_lib.SAMPLE_GUMBEL, gumbel=urows, gumbel_is_uniform=True,
                         sample_from=_lib.FROM_RECON, want_score=True, want_recon=True)
    want_score = torch.exp(recon).max(dim=1).values.clamp(0, 1)
    assert (out["score"].cpu() - want_score).abs().max() <= 1e-6
    want = O.log_sample_categorical(recon, u.permute(0, 2, 1), return_index=True)
    H.assert_tokens_match(out["x_prev"].cpu().numpy(), want.numpy(), O.near_ties(recon, u.permute(0, 2, 1)).numpy(), "cand")
    # production noise: Philox thinned race == Philox exhaustive on the same stream
    a = ops.fused_step(lc.to(DEV), lu.to(DEV), x_t.to(DEV), t.to(DEV), table, guidance_scale=2.0,
                       sample_mode=_lib.SAMPLE_PHILOX, seed=3, offset=1, sample_from=_lib.FROM_RECON)["x_prev"]
    b = ops.fused_step(lc.to(DEV), lu.to(DEV), x_t.to(DEV), t.to(DEV), table, guidance_scale=2.0,
                       sample_mode=_lib.SAMPLE_PHILOX_EXACT, seed=3, offset=1, sample_from=_lib.FROM_RECON)["x_prev"]
    assert torch.equal(a, b) and int(a.max()) < K


def test_purity_select_properties():
    """d3pm_purity_select at a production shape (N = 4096): exactly n_reveal[b] [MASK] positions are revealed, nothing else
    changes, the choice equals topk(w / q) computed by torch on the same noise, and the Philox path is deterministic."""
    B, N, K = 4, 4096, 4096
    g = torch.Generator().manual_seed(0)
    x_t = torch.where(torch.rand(B, N, generator=g) < 0.6, torch.full((B, N), K), torch.randint(0, K, (B, N), generator=g))
    cand = torch.randint(0, K, (B, N), generator=g)
    score = torch.rand(B, N, generator=g)
    expo = torch.empty(B, N).exponential_(1, generator=g)
    n = torch.tensor([0, 1, 11, 700], dtype=torch.int32)
    out, rev = ops.purity_select(x_t.to(DEV), cand.to(DEV), score.to(DEV), n.to(DEV), K, expo=expo.to(DEV))
    out, rev = out.cpu(), rev.cpu()
    w = score / (score.max(dim=1, keepdim=True).values + 1e-10)
    w[x_t != K] = 0
    for b in range(B):
        sel = torch.topk(w[b] / expo[b], int(n[b])).indices if int(n[b]) else torch.zeros(0, dtype=torch.long)
        want = x_t[b].clone()
        want[sel] = cand[b][sel]
        assert torch.equal(out[b], want)
        assert int(rev[b]) == int(n[b])
    a1, _ = ops.purity_select(x_t.to(DEV), cand.to(DEV), score.to(DEV), n.to(DEV), K, seed=5, offset=9)
    a2, _ = ops.purity_select(x_t.to(DEV), cand.to(DEV), score.to(DEV), n.to(DEV), K, seed=5, offset=9)
    a3, r3 = ops.purity_select(x_t.to(DEV), cand.to(DEV), None, n.to(DEV), K, seed=6, offset=9)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)
    assert r3.cpu().tolist() == n.tolist()
    assert ((a3.cpu() != x_t).sum(1) <= n).all() and bool((x_t[a3.cpu() != x_t] == K).all())


def test_sample_chain_with_purity_prior():
    """Whole chain with prior_rule = 2: n_sample[t] tokens are revealed per step, nothing stays masked, reproducible."""
    K, T, N, B = 64, 100, 1024, 2
    g = torch.Generator().manual_seed(1)
    lc, lu = torch.randn(B, N, K, generator=g).to(DEV), torch.randn(B, N, K, generator=g).to(DEV)
    m = d3pm_b200.FusedDiffusionTransformer(transformer=StubDenoiser(K, lc, lu), diffusion_step=T, alpha_init_type="alpha1",
                                            guidance_scale=2.0, content_seq_len=N).to(DEV)
    m.prior_rule = 2
    cond, cf = torch.ones(B, 1, 512, device=DEV), torch.zeros(B, 1, 512, device=DEV)
    a = m.manual_seed(3).sample(["x"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    b = m.manual_seed(3).sample(["x"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    assert torch.equal(a, b) and not (a == K).any()
    # one step of the chain: exactly n_sample[t] reveals per video
    x = torch.full((B, N), K, dtype=torch.int64, device=DEV)
    t = torch.full((B,), 57, dtype=torch.int64, device=DEV)
    y, sampled = m.p_sample_tokens_purity(x, cond, cf, t, [0] * B, m.n_sample[57])
    assert sampled == [m.n_sample[57]] * B and ((y != K).sum(1) == m.n_sample[57]).all()


def test_candidate_draw_and_purity_on_the_stream_kernel():
    """The purity-prior candidate draw (x0 ~ p(x0 | x_t), D3PM_FROM_RECON) and the purity score on the persistent stream
    kernel: same candidates as its exhaustive mode and as the one-CTA-per-row kernel, scores equal to 1e-5."""
    from d3pm_b200 import _lib, ops
    from oracle import d3pm_oracle as O
    T, K, B, N = 100, 4096, 2, 1500
    dev = "cuda:0"
    table = ops.build_coef_table(O.pack_schedule(O.make_schedule(T, K)).to(dev), T, K)
    g = torch.Generator(device=dev).manual_seed(2)
    lc, lu = torch.randn(B, N, K, device=dev, generator=g) * 2, torch.randn(B, N, K, device=dev, generator=g) * 2
    x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < 0.6, torch.full((B, N), K, device=dev),
                      torch.randint(0, K, (B, N), device=dev, generator=g))
    t = torch.tensor([70, 20], device=dev)
    kw = dict(guidance_scale=2.0, seed=8, offset=3, sample_from=_lib.FROM_RECON, want_score=True)
    outs = {}
    for name, mode, kern in (("stream", _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM), ("stream_exact", _lib.SAMPLE_PHILOX_EXACT, _lib.KERNEL_STREAM),
                             ("rows", _lib.SAMPLE_PHILOX, _lib.KERNEL_ROWS), ("rows_exact", _lib.SAMPLE_PHILOX_EXACT, _lib.KERNEL_ROWS)):
        outs[name] = ops.fused_step(lc, lu, x_t, t, table, sample_mode=mode, kernel=kern, **kw)
    torch.cuda.synchronize()
    assert torch.equal(outs["stream"]["x_prev"], outs["stream_exact"]["x_prev"])
    assert torch.equal(outs["rows"]["x_prev"], outs["rows_exact"]["x_prev"])
    assert (outs["stream"]["x_prev"] != outs["rows"]["x_prev"]).float().mean() < 1e-3   # p may differ in the last bit
    torch.testing.assert_close(outs["stream"]["score"], outs["rows"]["score"], rtol=1e-5, atol=0)
    assert int(outs["stream"]["x_prev"].max()) < K        # a candidate is never [MASK]

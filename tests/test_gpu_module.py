"""`FusedDiffusionTransformer` (the drop-in class) against the reference's golden outputs.  Needs a B200."""
import types

import numpy as np
import pytest
import torch

import d3pm_b200
from d3pm_b200 import ops
from oracle import d3pm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class StubDenoiser(torch.nn.Module):
    """Returns fixed logits the way Text2ImageTransformer does: a [B,K,N] permuted view of [B,N,K]
    (transformer_utils.py:442-443); cond[...,0] > 0.5 selects the conditional tensor."""

    def __init__(self, K, lc, lu):
        super().__init__()
        self.content_emb = types.SimpleNamespace(num_embed=K + 1)
        self.to_logits = torch.nn.Sequential(torch.nn.Identity(), torch.nn.Linear(1, 1))
        self.lc, self.lu, self.calls = lc, lu, 0

    def forward(self, x_t, cond, t):
        self.calls += 1
        assert x_t.dtype == torch.int64 and x_t.dim() == 2
        return (self.lc if float(cond.flatten()[0]) > 0.5 else self.lu).permute(0, 2, 1)


def _model(fx, T=None, s=None):
    K = int(fx["K"])
    T = int(fx["T"]) if T is None else T
    lc, lu = torch.from_numpy(fx["logits_c"]).to(DEV), torch.from_numpy(fx["logits_u"]).to(DEV)
    gs = float(fx["guidance_scale"]) if s is None else s
    m = d3pm_b200.FusedDiffusionTransformer(transformer=StubDenoiser(K, lc, lu), diffusion_step=T,
                                            alpha_init_type="alpha1", guidance_scale=gs,
                                            content_seq_len=lc.shape[1]).to(DEV)
    B = lc.shape[0]
    return m, torch.ones(B, 1, 512, device=DEV), torch.zeros(B, 1, 512, device=DEV)


@pytest.mark.parametrize("name", ["k64_t50", "k64_t0", "k64_pert_s5_stress", "k4096_t50", "k4096_t0_stress"])
def test_method_surface_against_golden(name):
    fx = H.load(f"{H.GOLDEN}/step_{name}.npz")
    K = int(fx["K"])
    m, cond, cf = _model(fx)
    x_t, t = torch.from_numpy(fx["x_t"]).to(DEV), torch.from_numpy(fx["t"]).to(DEV)
    # the caller hands over a log one-hot in the reference's own (index_to_log_onehot) layout
    log_x = O.index_to_log_onehot(x_t.cpu(), K + 1).to(DEV)
    post, recon = m.p_pred(log_x, cond, cf, t)
    assert post.shape == recon.shape == (x_t.shape[0], K + 1, x_t.shape[1])
    want_post, want_recon = torch.from_numpy(fx["post"]).permute(0, 2, 1), torch.from_numpy(fx["recon"]).permute(0, 2, 1)
    assert (post.cpu() - want_post).abs().max() <= H.POST_TOL
    assert (recon.cpu() - want_recon).abs().max() <= H.POST_TOL
    assert (m.cf_predict_start(log_x, cond, cf, t).cpu() - want_recon).abs().max() <= H.POST_TOL
    # q_posterior as a stand-alone operator, fed the reference-layout (class-major contiguous) recon
    qp = m.q_posterior(want_recon.contiguous().to(DEV), log_x, t)
    assert (qp.cpu() - want_post).abs().max() <= H.POST_TOL
    # p_sample with the reference's uniform tensor injected where it calls torch.rand_like
    u = torch.from_numpy(fx["uniform"]).permute(0, 2, 1).contiguous()
    m.inject_uniform = lambda shape, dev: u.to(dev)
    out, sampled = m.p_sample(log_x, cond, cf, t, [0] * x_t.shape[0], 10)
    assert sampled == [1024] * x_t.shape[0] and out.shape == post.shape
    tok = out.argmax(1).cpu()
    H.assert_tokens_match(tok.numpy(), fx["x_prev"], fx["near_tie"], name)
    assert torch.equal(out.cpu(), O.index_to_log_onehot(tok, K + 1))
    # log_sample_categorical alone
    ls = m.log_sample_categorical(want_post.contiguous().to(DEV))
    H.assert_tokens_match(ls.argmax(1).cpu().numpy(), fx["x_prev"], fx["near_tie"], name + " lsc")
    m.check_status()


def test_predict_start_guidance_off():
    fx = H.load(f"{H.GOLDEN}/step_k64_t50_noguid.npz")
    K = int(fx["K"])
    m, cond, cf = _model(fx, s=2.0)
    x_t, t = torch.from_numpy(fx["x_t"]).to(DEV), torch.from_numpy(fx["t"]).to(DEV)
    log_x = ops.as_logical(ops.tokens_to_log_onehot_rows(x_t, K + 1), K + 1)
    recon = m.predict_start(log_x, cond, t)
    assert (recon.cpu() - torch.from_numpy(fx["recon"]).permute(0, 2, 1)).abs().max() <= H.POST_TOL
    m.guidance_scale = 1.0  # the reference crashes here (:242-243); we return predict_start's result
    assert (m.cf_predict_start(log_x, cond, cf, t).cpu() - recon.cpu()).abs().max() == 0


def test_sample_loop_against_reference_trace():
    """The reference's complete sample() chain (T=10, uniforms injected per step)."""
    fx = H.load(f"{H.GOLDEN}/sample_loop_k64_T10.npz")
    K, T = int(fx["K"]), int(fx["T"])
    m, cond, cf = _model(fx, T=T)
    B = cond.shape[0]
    us = iter([torch.from_numpy(u).permute(0, 2, 1).contiguous() for u in fx["uniforms"]])
    m.inject_uniform = lambda shape, dev: next(us).to(dev)
    res = m.sample(["a"] * B, None, cond, cf, content_token=None, filter_ratio=0)
    assert m.transformer.calls == 2 * T
    assert res["content_token"].dtype == torch.int64
    assert np.array_equal(res["content_token"].cpu().numpy(), fx["content_token"])
    m.inject_uniform = None
    a = m.manual_seed(5).sample(["a"] * B, None, cond, cf, filter_ratio=0, return_logits=True)
    b = m.manual_seed(5).sample(["a"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    c = m.manual_seed(6).sample(["a"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    assert torch.equal(a["content_token"], b) and not torch.equal(b, c)
    assert not (b == K).any() and a["logits"].shape == (B, K + 1, b.shape[1])


def test_unbuilt_paths_fail_loudly():
    fx = H.load(f"{H.GOLDEN}/step_k64_t50.npz")
    m, cond, cf = _model(fx)
    K = int(fx["K"])
    x_t, t = torch.from_numpy(fx["x_t"]).to(DEV), torch.from_numpy(fx["t"]).to(DEV)
    log_x = ops.as_logical(ops.tokens_to_log_onehot_rows(x_t, K + 1), K + 1)
    with pytest.raises(NotImplementedError):
        m.sample(["a", "b"], None, cond, cf, filter_ratio=0.5)
    with pytest.raises(NotImplementedError):
        m.q_posterior(log_x.clone().requires_grad_(True), log_x, t)
    bad_t = torch.full_like(t, 100)
    m.p_sample_tokens(x_t, cond, cf, bad_t)
    with pytest.raises(AssertionError):
        m.check_status()


@pytest.mark.parametrize("B,guidance", [(8, True), (3, True), (4, False)])
def test_host_step_equals_the_device_resident_step(B, guidance):
    """`ops.HostStep` (pinned host logits in, host tokens out - the call `bench.py` times as e2e) sends the videos in
    chunks and runs the step of one chunk while the next is on the bus: the tokens are those of one launch over
    device-resident inputs."""
    from d3pm_b200 import _lib, ops
    from oracle import d3pm_oracle as O
    T, K, N = 100, 4096, 1024
    dev = "cuda:0"
    table = ops.build_coef_table(O.pack_schedule(O.make_schedule(T, K)).to(dev), T, K)
    g = torch.Generator().manual_seed(B)
    lc = torch.randn(B, N, K, generator=g).pin_memory()
    lu = torch.randn(B, N, K, generator=g).pin_memory() if guidance else None
    x_t = torch.randint(0, K + 1, (B, N), generator=g).pin_memory()
    t = torch.randint(0, T, (B,), generator=g).pin_memory()
    host = ops.HostStep(B, N, K, table, guidance=guidance)
    got = host(lc, lu, x_t, t, guidance_scale=2.0, seed=5, offset=9, row_offset=77).clone()
    want = ops.fused_step(lc.to(dev), lu.to(dev) if guidance else None, x_t.to(dev), t.to(dev), table, guidance_scale=2.0,
                          sample_mode=_lib.SAMPLE_PHILOX, seed=5, offset=9, row_offset=77)["x_prev"]
    assert torch.equal(got, want.cpu())
    again = host(lc, lu, x_t, t, guidance_scale=2.0, seed=5, offset=9, row_offset=77)
    assert torch.equal(again, got)


def test_host_head_step_equals_the_device_resident_head_step():
    """The host entry of the fused head (`d3pm_host_head_step_run`: pinned host HIDDEN STATES in, host tokens out; 64x fewer
    bytes on the bus than the logits) against `head.head_step` on device-resident inputs, and a bad timestep is reported
    through the status word of the call."""
    from d3pm_b200 import _lib, head, ops
    from oracle import d3pm_oracle as O
    T, K, N, B, D = 100, 4096, 1024, 8, 64
    dev = "cuda:0"
    table = ops.build_coef_table(O.pack_schedule(O.make_schedule(T, K)).to(dev), T, K)
    g = torch.Generator().manual_seed(21)
    tl = torch.nn.Sequential(torch.nn.LayerNorm(D), torch.nn.Linear(D, K)).to(dev)
    hw = head.HeadWeights.from_module(tl)
    assert hw.valid
    hc, hu = torch.randn(B, N, D, generator=g).pin_memory(), torch.randn(B, N, D, generator=g).pin_memory()
    x_t = torch.randint(0, K + 1, (B, N), generator=g).pin_memory()
    t = torch.randint(0, T, (B,), generator=g).pin_memory()
    host = ops.HostStep(B, N, K, table, guidance=True, hidden_dim=D)
    got = host.head(hw, hc, hu, x_t, t, guidance_scale=2.0, seed=3, offset=4, row_offset=11).clone()
    want = head.head_step(hw, hc.to(dev), hu.to(dev), x_t.to(dev), t.to(dev), table, guidance_scale=2.0, seed=3, offset=4, row_offset=11)
    assert torch.equal(got, want.cpu())
    assert host.last_status & (_lib.STATUS_BAD_T | _lib.STATUS_BAD_TOKEN) == 0
    assert host.h2d_bytes == 2 * B * N * D * 4 + B * N * 8 + B * 8 and host.d2h_bytes == B * N * 8
    t_bad = t.clone()
    t_bad[2] = T + 5
    host.head(hw, hc, hu, x_t, t_bad.pin_memory(), guidance_scale=2.0, seed=3, offset=4)
    assert host.last_status & _lib.STATUS_BAD_T
    host.close()
    with pytest.raises(ops.D3PMError):
        ops.HostStep(B, N, K, table, guidance=True)(hc, hu, x_t, t, guidance_scale=2.0, seed=1, offset=1)  # wrong width


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_16bit_denoiser_logits_are_stepped_in_place(dtype):
    """A denoiser that returns half-precision logits (autocast): `p_sample_tokens` hands them to the stream kernel as they
    are (no `.float()` pass) and draws the tokens of the fp32 path on the up-cast logits; the methods that produce rows
    (`p_pred`) still work through the cast."""
    K, B, N, Tn = 4096, 2, 1024, 100
    g = torch.Generator(device=DEV).manual_seed(3)
    lc, lu = torch.randn(B, N, K, device=DEV, generator=g).to(dtype), torch.randn(B, N, K, device=DEV, generator=g).to(dtype)
    cond, cf = torch.ones(B, 1, 512, device=DEV), torch.zeros(B, 1, 512, device=DEV)

    def model(a, b):
        return d3pm_b200.FusedDiffusionTransformer(transformer=StubDenoiser(K, a, b), diffusion_step=Tn, alpha_init_type="alpha1",
                                                   guidance_scale=2.0, content_seq_len=N).to(DEV)

    m16, m32 = model(lc, lu), model(lc.float(), lu.float())
    x_t = torch.randint(0, K + 1, (B, N), device=DEV, generator=g)
    t = torch.full((B,), 40, dtype=torch.long, device=DEV)
    seen = []
    real = ops.fused_step
    try:
        ops.fused_step = lambda *a, **k: (seen.append(a[0].dtype), real(*a, **k))[1]
        a = m16.manual_seed(9).p_sample_tokens(x_t, cond, cf, t)
    finally:
        ops.fused_step = real
    b = m32.manual_seed(9).p_sample_tokens(x_t, cond, cf, t)
    assert seen == [dtype] and torch.equal(a, b)
    log_x = ops.as_logical(ops.tokens_to_log_onehot_rows(x_t, K + 1), K + 1)
    pa, pb = m16.p_pred(log_x, cond, cf, t), m32.p_pred(log_x, cond, cf, t)
    assert torch.equal(pa[0], pb[0]) and torch.equal(pa[1], pb[1])
    m16.check_status()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_host_step_with_16bit_host_logits(dtype):
    """Host logits in half precision: half the bytes on the bus, the tokens of the device-resident 16-bit step."""
    from d3pm_b200 import _lib
    K, B, N, Tn = 4096, 4, 1024, 100
    g = torch.Generator().manual_seed(12)
    lc, lu = torch.randn(B, N, K, generator=g).to(dtype), torch.randn(B, N, K, generator=g).to(dtype)
    x_t = torch.randint(0, K + 1, (B, N), generator=g)
    t = torch.full((B,), 30, dtype=torch.long)
    table = ops.build_coef_table(O.pack_schedule(O.make_schedule(Tn, K)).to(DEV), Tn, K)
    hs = ops.HostStep(B, N, K, table, guidance=True, logits_dtype=dtype)
    ref32 = ops.HostStep(B, N, K, table, guidance=True)
    assert hs.h2d_bytes - B * N * 8 - B * 8 == (ref32.h2d_bytes - B * N * 8 - B * 8) // 2
    got = hs(lc.pin_memory(), lu.pin_memory(), x_t.pin_memory(), t.pin_memory(), guidance_scale=2.0, seed=4, offset=9).clone()
    want = ops.fused_step(lc.to(DEV), lu.to(DEV), x_t.to(DEV), t.to(DEV), table, guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX,
                          seed=4, offset=9)["x_prev"].cpu()
    assert torch.equal(got, want) and hs.last_status & 3 == 0
    with pytest.raises(Exception):
        hs(lc.float().pin_memory(), lu.float().pin_memory(), x_t, t, guidance_scale=2.0, seed=4, offset=9)
    hs.close(), ref32.close()

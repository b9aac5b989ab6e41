"""Fused denoiser head + reverse step against torch fp32, the oracle and the unfused
CUDA path.  Needs a B200 (tcgen05)."""
import copy
import math

import numpy as np
import pytest
import torch

import d3pm_b200
from d3pm_b200 import _lib, head, ops
from d3pm_b200._lib import D3PMError
from oracle import d3pm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
D, T = 64, 100


def _head(K, seed, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    ln, lin = torch.nn.LayerNorm(D), torch.nn.Linear(D, K)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.1 * torch.randn(D, generator=g))
        ln.bias.copy_(0.1 * torch.randn(D, generator=g))
        lin.weight.copy_(torch.randn(K, D, generator=g) * (scale / 8 / math.sqrt(3)))
        lin.bias.copy_(torch.randn(K, generator=g) * 0.5)
    return torch.nn.Sequential(ln, lin)


def _inputs(B, N, K, tval, seed):
    g = torch.Generator().manual_seed(seed)
    hc = torch.randn(B, N, D, generator=g) * 2 + 0.3
    hu = torch.randn(B, N, D, generator=g)
    sched = O.make_schedule(T, K)
    t = tval.clone() if torch.is_tensor(tval) else torch.full((B,), tval, dtype=torch.long)
    p_mask = sched["log_cumprod_ct"][t].exp().view(B, 1)
    x_t = torch.where(torch.rand(B, N, generator=g) < p_mask, torch.full((B, N), K), torch.randint(0, K, (B, N), generator=g))
    return hc, hu, x_t, t, sched


@pytest.mark.parametrize("K,B,N,guid", [(4096, 2, 96, True), (4096, 1, 128, False), (2048, 3, 50, True), (1024, 1, 300, True)])
def test_combined_logits_against_torch(K, B, N, guid):
    """LOGITS mode = s * Linear(LayerNorm(h_c)) + (1 - s) * Linear(LayerNorm(h_u)) to fp32 accuracy (3xTF32 on tcgen05);
    row counts that are not multiples of the 128-row tile included."""
    tl = _head(K, 1)
    hc, hu, _, _, _ = _inputs(B, N, K, 50, 2)
    s = 2.0
    tl64 = copy.deepcopy(tl).double()
    with torch.no_grad():
        lc = tl64(hc.double())
        lu = tl64(hu.double()) if guid else None
    want = s * lc + (1 - s) * lu if guid else lc
    hw = head.HeadWeights.from_module(tl.to(DEV))
    assert hw.valid
    got = head.head_step(hw, hc.to(DEV), hu.to(DEV) if guid else None, None, None, None, guidance_scale=s, mode=_lib.HEAD_LOGITS)
    err = (got.cpu().double() - want).abs().max().item()
    assert err <= 5e-5, f"combined logits differ from float64 torch by {err} (|logit| up to {want.abs().max().item():.1f})"


@pytest.mark.parametrize("K,B,N,tval,guid", [(4096, 2, 96, 50, True), (4096, 2, 70, [0, 99], True), (1024, 2, 130, 7, True),
                                             (2048, 1, 200, 30, False)])
def test_tokens_against_oracle(K, B, N, tval, guid):
    """The fused kernel's tokens == the reference algorithm (oracle port, PyTorch CPU head + p_sample) fed the very uniforms
    the kernel's Philox stream draws, except at logged near-ties."""
    tl = _head(K, 3)
    tv = torch.tensor(tval) if isinstance(tval, list) else tval
    hc, hu, x_t, t, sched = _inputs(B, N, K, tv, 4)
    s = 2.0
    with torch.no_grad():
        lc = tl(hc).permute(0, 2, 1)
        lu = tl(hu).permute(0, 2, 1) if guid else None
    hw = head.HeadWeights.from_module(tl.to(DEV))
    table = ops.build_coef_table(O.pack_schedule(sched).to(DEV), T, K)
    status = ops.new_status(DEV)
    args = (hw, hc.to(DEV), hu.to(DEV) if guid else None, x_t.to(DEV), t.to(DEV), table)
    fused = head.head_step(*args, guidance_scale=s, seed=11, offset=5, row_offset=1000, status=status).cpu()
    ref = head.head_step(*args, guidance_scale=s, mode=_lib.HEAD_REFERENCE, seed=11, offset=5, row_offset=1000).cpu()
    u = ops.philox_uniform(B, N, K, seed=11, offset=5, row_offset=1000, device=DEV)[:, :, :K + 1].cpu().permute(0, 2, 1)
    out, post, _ = O.p_sample_step(sched, lc, lu, O.index_to_log_onehot(x_t, K + 1), t, s if guid else 0.0, u)
    want, ties = out.argmax(1).numpy(), O.near_ties(post, u).numpy()
    H.assert_tokens_match(fused.numpy(), want, ties, "fused head")
    H.assert_tokens_match(ref.numpy(), want, ties, "CUDA-core reference")
    assert int(status.item()) & 3 == 0
    assert (fused != x_t).any()  # the step does something


def test_fused_equals_unfused_at_scale():
    """16 k rows: tensor-core path == CUDA-core reference path == d3pm_fused_step on torch-made logits, same Philox stream."""
    K, B, N, s = 4096, 4, 4096, 2.0
    tl = _head(K, 5).to(DEV)
    hc, hu, x_t, t, sched = _inputs(B, N, K, torch.tensor([80, 50, 20, 0]), 6)
    hc, hu, x_t, t = hc.to(DEV), hu.to(DEV), x_t.to(DEV), t.to(DEV)
    table = ops.build_coef_table(O.pack_schedule(sched).to(DEV), T, K)
    hw = head.HeadWeights.from_module(tl)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        lc, lu = tl(hc), tl(hu)
    unf = ops.fused_step(lc, lu, x_t, t, table, guidance_scale=s, sample_mode=_lib.SAMPLE_PHILOX_EXACT, seed=2, offset=9,
                         want_gap=True, kernel=_lib.KERNEL_ROWS)
    fused = head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=s, seed=2, offset=9)
    ref = head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=s, seed=2, offset=9, mode=_lib.HEAD_REFERENCE)
    near = (unf["gap"] < H.NEAR_TIE_GAP).cpu().numpy()
    H.assert_tokens_match(fused.cpu().numpy(), unf["x_prev"].cpu().numpy(), near, "fused vs unfused")
    H.assert_tokens_match(ref.cpu().numpy(), unf["x_prev"].cpu().numpy(), near, "reference vs unfused")
    # the statistics pass in 3xTF32 instead of 1xTF32 (wider vs exact thinning thresholds): the very same tokens
    fused3 = head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=s, seed=2, offset=9, stats_1xtf32=False)
    assert torch.equal(fused3, fused)
    assert 0.0 < hw.stat_slack(s) < 0.5
    # a forced thinning failure (tiny c): every row goes through the redo kernel and the result is unchanged
    redo = head.head_step(hw, hc[:1], hu[:1], x_t[:1], t[:1], table, guidance_scale=s, seed=2, offset=9, thin_factor=1e-3)
    H.assert_tokens_match(redo.cpu().numpy(), unf["x_prev"][:1].cpu().numpy(), near[:1], "all rows redone")


class HeadDenoiser(torch.nn.Module):
    """Embedding + the reference's head; returns a `[B, K, N]` view of `[B, N, K]` like Text2ImageTransformer (:442-443)."""

    def __init__(self, K, N):
        super().__init__()
        self.content_emb = torch.nn.Embedding(K + 1, D)
        self.content_emb.num_embed = K + 1
        self.pos = torch.nn.Parameter(torch.randn(N, D) * 0.5)
        self.to_logits = _head(K, 7)

    def hidden_states(self, x_t, cond, t):
        return self.content_emb(x_t) + self.pos + cond.mean(-1, keepdim=True) + 0.01 * t[:, None, None]

    def forward(self, x_t, cond, t):
        return self.to_logits(self.hidden_states(x_t, cond, t)).permute(0, 2, 1)


def test_drop_in_class_with_fused_head():
    K, N, B = 1024, 256, 2
    torch.manual_seed(0)
    den = HeadDenoiser(K, N).to(DEV)
    m = d3pm_b200.FusedDiffusionTransformer(transformer=den, diffusion_step=T, alpha_init_type="alpha1", guidance_scale=2.0,
                                            content_seq_len=N).to(DEV)
    cond, cf = torch.randn(B, 1, 512, device=DEV), torch.zeros(B, 1, 512, device=DEV)
    x = torch.full((B, N), K, dtype=torch.int64, device=DEV)
    x[:, ::3] = 5
    t = torch.full((B,), 40, dtype=torch.int64, device=DEV)
    a = m.manual_seed(9).p_sample_tokens(x, cond, cf, t)
    m.enable_fused_head()
    assert m.fused_head_active
    b = m.manual_seed(9).p_sample_tokens(x, cond, cf, t)
    assert isinstance(m.transformer.to_logits, torch.nn.Sequential)  # the module tree is never touched
    assert (a != b).float().mean().item() <= 0.01 and (a != x).any()
    out = m.manual_seed(1).sample(["x"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    assert out.shape == (B, N) and not (out == K).any()
    m.check_status()
    # weights outside the no-clamp bound: the class keeps the unfused (exact) path
    with torch.no_grad():
        den.to_logits[-1].weight.mul_(50.0)
    assert not m.fused_head_active
    c = m.manual_seed(9).p_sample_tokens(x, cond, cf, t)
    m.enable_fused_head(False)
    d = m.manual_seed(9).p_sample_tokens(x, cond, cf, t)
    assert torch.equal(c, d)


def test_unsupported_heads_fail_loudly():
    with pytest.raises(D3PMError):
        head.HeadWeights.from_module(torch.nn.Sequential(torch.nn.LayerNorm(32), torch.nn.Linear(32, 4096)).to(DEV))
    with pytest.raises(D3PMError):
        head.HeadWeights.from_module(torch.nn.Sequential(torch.nn.LayerNorm(64), torch.nn.Linear(64, 1000)).to(DEV))
    with pytest.raises(D3PMError):
        head.HeadWeights.from_module(torch.nn.Sequential(torch.nn.Identity(), torch.nn.Linear(64, 4096)).to(DEV))
    hw = head.HeadWeights.from_module(_head(1024, 1).to(DEV))
    with pytest.raises(D3PMError):
        head.head_step(hw, torch.zeros(1, 8, 64), None, None, None, None, guidance_scale=1.0, mode=_lib.HEAD_LOGITS)  # CPU tensor

"""compute-sanitizer is closed on this pool, so the newer kernels are fenced by hand: NaN guard bands around every input
(an out-of-window read poisons the result) and sentinels around every output (an out-of-window write clobbers them), at
row counts that leave the last tile / group partially filled.  Needs a B200."""
import pytest
import torch

import d3pm_b200
from d3pm_b200 import _lib, head, ops, train
from oracle import d3pm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
T = 100


def _fenced(shape, dtype, fill, pad=256):
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((pad + n + pad,), fill, dtype=dtype, device=DEV)
    return buf, buf[pad:pad + n].view(*shape)


def _intact(buf, fill, pad=256):
    if isinstance(fill, float) and fill != fill:
        return bool(torch.isnan(buf[:pad]).all() and torch.isnan(buf[-pad:]).all())
    return bool((buf[:pad] == fill).all() and (buf[-pad:] == fill).all())


def _table(K):
    return ops.build_coef_table(O.pack_schedule(O.make_schedule(T, K)).to(DEV), T, K)


def test_head_step_stays_inside_its_buffers():
    K, B, N, D = 4096, 3, 171, 64  # 513 rows: four full tiles and one row in the fifth
    g = torch.Generator(device=DEV).manual_seed(1)
    tl = torch.nn.Sequential(torch.nn.LayerNorm(D), torch.nn.Linear(D, K)).to(DEV)
    hw = head.HeadWeights.from_module(tl)
    nan = float("nan")
    bc, hc = _fenced((B, N, D), torch.float32, nan)
    bu, hu = _fenced((B, N, D), torch.float32, nan)
    hc.copy_(torch.randn(B, N, D, device=DEV, generator=g))
    hu.copy_(torch.randn(B, N, D, device=DEV, generator=g))
    bx, x_t = _fenced((B, N), torch.int64, -1)
    x_t.copy_(torch.randint(0, K + 1, (B, N), device=DEV, generator=g))
    bt, t = _fenced((B,), torch.int64, -1)
    t.copy_(torch.tensor([0, 50, 99], device=DEV))
    bo, x_prev = _fenced((B, N), torch.int64, -7)
    br, redo = _fenced((B * N,), torch.int32, -9)
    bcnt, cnt = _fenced((1,), torch.int32, -9)
    cnt.zero_()
    status = ops.new_status(DEV)
    table = _table(K)
    # thin_factor tiny: every row also goes through the redo list and the redo kernel
    for thin in (0.0, 1e-3):
        head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=2.0, seed=4, offset=2, status=status, x_prev_out=x_prev,
                       thin_factor=thin, scratch=(redo, cnt))
        torch.cuda.synchronize()
        assert _intact(bo, -7) and _intact(br, -9) and _intact(bcnt, -9)
        assert int(x_prev.min()) >= 0 and int(x_prev.max()) <= K and int(status.item()) & 3 == 0
        assert int(cnt.item()) <= B * N
    dense = head.head_step(hw, hc.clone(), hu.clone(), x_t.clone(), t.clone(), table, guidance_scale=2.0, seed=4, offset=2)
    redone = x_prev.clone()
    head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=2.0, seed=4, offset=2, x_prev_out=x_prev, scratch=(redo, cnt))
    assert torch.equal(dense, x_prev)
    assert (redone != x_prev).float().mean().item() <= 0.002  # fp32 redo path vs tensor-core path: near-ties only
    bl, _ = _fenced((1,), torch.float32, 0.0)
    logits = head.head_step(hw, hc, hu, None, None, None, guidance_scale=2.0, mode=_lib.HEAD_LOGITS)
    assert torch.isfinite(logits).all()


def test_purity_select_stays_inside_its_buffers():
    K, B, N = 4096, 3, 1000  # not a power of two: the sort pads to 1024 keys
    g = torch.Generator(device=DEV).manual_seed(2)
    bx, x_t = _fenced((B, N), torch.int64, -1)
    x_t.copy_(torch.where(torch.rand(B, N, device=DEV, generator=g) < 0.5, torch.full((B, N), K, device=DEV),
                          torch.randint(0, K, (B, N), device=DEV, generator=g)))
    bc, cand = _fenced((B, N), torch.int64, -1)
    cand.copy_(torch.randint(0, K, (B, N), device=DEV, generator=g))
    bs, score = _fenced((B, N), torch.float32, float("nan"))
    score.copy_(torch.rand(B, N, device=DEV, generator=g))
    n = torch.tensor([5, 0, 300], dtype=torch.int32, device=DEV)
    out, rev = ops.purity_select(x_t, cand, score, n, K, seed=1, offset=1)
    assert rev.tolist() == [5, 0, 300]
    changed = out != x_t
    assert changed.sum(1).tolist() == [5, 0, 300] and bool((x_t[changed] == K).all())
    assert torch.equal(out[changed], cand[changed])


def test_train_stream_stays_inside_its_buffers():
    K, B, N = 1024, 3, 700  # 2100 rows: the persistent kernel, last groups partially filled
    g = torch.Generator(device=DEV).manual_seed(3)
    bl, logits = _fenced((B, N, K), torch.float32, float("nan"))
    logits.copy_(torch.randn(B, N, K, device=DEV, generator=g))
    x0 = torch.randint(0, K, (B, N), device=DEV, generator=g)
    x_t = torch.where(torch.rand(B, N, device=DEV, generator=g) < 0.5, torch.full((B, N), K, device=DEV), x0)
    t = torch.tensor([0, 40, 99], device=DEV)
    w = torch.ones(B, device=DEV)
    out = train._train_rows(logits, K, x0, x_t, t, _table(K), (1, 1), backward=2, w_main=w, w_aux=w, want_recon=True)
    for name in ("grad", "tok_main", "tok_aux"):
        assert torch.isfinite(out[name]).all(), name
    assert int(out["x0_recon"].max()) < K and int(out["xtm1_recon"].max()) <= K
    dense = train._train_rows(logits.clone(), K, x0, x_t, t, _table(K), (1, 1), backward=2, w_main=w, w_aux=w, want_recon=True)
    assert torch.equal(dense["grad"], out["grad"]) and torch.equal(dense["tok_main"], out["tok_main"])


def test_q_sample_tokens_kernel_stays_inside_its_buffers():
    """d3pm_q_sample_tokens at a row count that leaves the last 8-warp CTA partly empty: sentinels around the token output,
    out-of-range fences around x_0 / t (any use of them would flag the status word or change the [MASK] rate)."""
    K, B, N = 1024, 3, 333            # 999 rows: 124 full CTAs + 7 warps
    bx, x0 = _fenced((B, N), torch.int64, -5)
    bt, t = _fenced((B,), torch.int64, 10 ** 9)
    bo, out = _fenced((B, N), torch.int64, -7)
    g = torch.Generator(device=DEV).manual_seed(3)
    x0.copy_(torch.randint(0, K, (B, N), device=DEV, generator=g))
    t.copy_(torch.tensor([99, 50, 0], device=DEV))
    sched8 = torch.zeros(8, T + 1, device=DEV)
    sched = O.make_schedule(T, K)
    for i, n in enumerate(("log_at", "log_bt", "log_ct", "log_1_min_ct", "log_cumprod_at", "log_cumprod_bt", "log_cumprod_ct",
                           "log_1_min_cumprod_ct")):
        sched8[i, : sched[n].numel()] = sched[n].to(DEV)
    status = ops.new_status(DEV)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_q_sample_tokens(x0.data_ptr(), t.data_ptr(), sched8.data_ptr(), B, N, K, T, 11, 2, 0, out.data_ptr(),
                                        status.data_ptr(), torch.cuda.current_stream().cuda_stream), "d3pm_q_sample_tokens")
    torch.cuda.synchronize()
    assert _intact(bo, -7) and _intact(bx, -5) and _intact(bt, 10 ** 9)
    assert int(status.item()) == 0
    assert int(out.min()) >= 0 and int(out.max()) <= K
    assert (out[0] == K).float().mean() > 0.97 and (out[2] == x0[2]).float().mean() > 0.99   # t = 99: masked, t = 0: kept

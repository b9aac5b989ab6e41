"""The persistent TMA-pipelined production kernel (D3PM_KERNEL_STREAM) against its own exhaustive mode,
the one-CTA-per-row kernel and the oracle fed the very noise the kernel drew.  Needs a B200."""
import numpy as np
import pytest
import torch

from d3pm_b200 import _lib, ops
from oracle import d3pm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
T = 100


def _table(K):
    return ops.build_coef_table(O.pack_schedule(O.make_schedule(T, K)).to(DEV), T, K)


def _run(lc, lu, x_t, t, K, mode, kernel, s=2.0, **kw):
    out = ops.fused_step(lc, lu, x_t, t, _table(K), guidance_scale=s, sample_mode=mode, kernel=kernel, **kw)
    torch.cuda.synchronize()
    return out["x_prev"]


CASES = [
    # B, N, K, t, guidance, logit scale
    (2, 1024, 4096, 50, 2.0, 1.0),
    (2, 1024, 4096, 0, 2.0, 1.0),
    (2, 1024, 4096, 99, 2.0, 1.0),
    (3, 700, 4096, [0, 42, 99], 3.0, 4.0),   # ragged row count, per-video t, peaked logits
    (2, 1500, 2048, 25, None, 1.0),          # guidance off, the UCF job's 2048-code book
    (1, 2100, 1024, 60, 2.0, 1.0),
    (1, 37, 4096, 50, 2.0, 1.0),             # fewer rows than groups
    (2, 900, 4096, 30, 2.0, 12.0),           # logit range > 70 - ln K: every row takes the general (clamping) path
    (2, 900, 4096, 0, 2.0, -1.0),            # scale < 0: a few rows spiked by +150 / +90 (mixed fast / general rows)
    # every group shape of the kernel (threads per row = K / 32, or K / 64 without guidance): <4,16,off> 8 groups of 64,
    # <2,8,on> 8 groups of 64, <2,16,off> / <1,8,on> / <1,8,off> 16 one-warp groups
    (2, 1100, 4096, 40, None, 1.0),
    (2, 700, 4096, [0, 99], None, 12.0),
    (2, 1300, 2048, 60, 2.0, 1.0),
    (2, 900, 2048, [5, 77], 2.0, 12.0),
    (2, 900, 2048, 0, 2.0, -1.0),
    (1, 2100, 1024, 33, None, 1.0),
    (3, 500, 1024, [0, 50, 99], 2.0, -1.0),
]


@pytest.mark.parametrize("B,N,K,tval,s,scale", CASES)
def test_stream_kernel_parity(B, N, K, tval, s, scale):
    sched = O.make_schedule(T, K)
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + N)
    spiked = scale < 0
    scale = abs(scale)
    lc = torch.randn(B, N, K, device=DEV, generator=g) * scale
    lu = None if s is None else torch.randn(B, N, K, device=DEV, generator=g) * scale
    if spiked:  # every 7th row gets a dominant class in each tensor (different classes)
        lc[:, ::7, 5] += 150.0
        lu[:, ::7, 9] += 90.0
    t = (torch.tensor(tval) if isinstance(tval, list) else torch.full((B,), tval)).long().to(DEV)
    p_mask = sched["log_cumprod_ct"][t.cpu()].exp().view(B, 1).to(DEV)
    x_t = torch.where(torch.rand(B, N, device=DEV, generator=g) < p_mask, torch.full((B, N), K, device=DEV),
                      torch.randint(0, K, (B, N), device=DEV, generator=g))
    kw = dict(s=0.0 if s is None else s, seed=99, offset=3, row_offset=12345)
    status = ops.new_status(DEV)
    thin = _run(lc, lu, x_t, t, K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, status=status, **kw)
    assert int(status.item()) & (_lib.STATUS_BAD_T | _lib.STATUS_BAD_TOKEN) == 0
    exact = _run(lc, lu, x_t, t, K, _lib.SAMPLE_PHILOX_EXACT, _lib.KERNEL_STREAM, **kw)
    assert torch.equal(thin, exact)                       # thinning never changes the draw
    forced = _run(lc, lu, x_t, t, K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, thin_factor=1e-3, status=status, **kw)
    assert torch.equal(forced, thin) and int(status.item()) & _lib.STATUS_FALLBACK
    tiny = _run(lc, lu, x_t, t, K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, thin_factor=1.0, **kw)  # many redone rows
    assert torch.equal(tiny, thin)
    assert int(thin.min()) >= 0 and int(thin.max()) <= K

    # against the oracle with the dumped uniforms (a sample of rows: the CPU oracle is slow)
    rows = torch.arange(0, N, max(1, N // 24))
    u = ops.philox_uniform(B, N, K, seed=99, offset=3, row_offset=12345, device=DEV)[:, rows, :K + 1].cpu().permute(0, 2, 1)
    lc_s = lc[:, rows].cpu().permute(0, 2, 1)
    lu_s = None if lu is None else lu[:, rows].cpu().permute(0, 2, 1)
    out_o, post_o, _ = O.p_sample_step(sched, lc_s, lu_s, O.index_to_log_onehot(x_t[:, rows].cpu(), K + 1), t.cpu(),
                                       0.0 if s is None else s, u)
    ties = O.near_ties(post_o, u).numpy()
    H.assert_tokens_match(thin[:, rows].cpu().numpy(), out_o.argmax(1).numpy(), ties, "stream vs oracle")
    # the posterior log-prob of the drawn class as the stream kernel computed it, on each of its three scoring paths
    # (batched survivor scoring, second thinned attempt, exhaustive), against the oracle's posterior row
    for tf in (0.0, 1.0, 1e-3):
        o = ops.fused_step(lc, lu, x_t, t, _table(K), guidance_scale=kw["s"], sample_mode=_lib.SAMPLE_PHILOX,
                           kernel=_lib.KERNEL_STREAM, seed=99, offset=3, row_offset=12345, thin_factor=tf, want_winner_post=True)
        assert torch.equal(o["x_prev"], thin)
        want = post_o.gather(1, thin[:, rows].cpu().unsqueeze(1)).squeeze(1)
        assert (o["winner_post"][:, rows].cpu() - want).abs().max().item() <= H.POST_TOL, f"thin_factor {tf}"
    # and against the other kernel (same noise definition; p may differ in the last bit)
    rowsk = _run(lc, lu, x_t, t, K, _lib.SAMPLE_PHILOX, _lib.KERNEL_ROWS, **kw)
    H.assert_tokens_match(rowsk[:, rows].cpu().numpy(), out_o.argmax(1).numpy(), ties, "rows vs oracle")
    assert (rowsk != thin).float().mean() < 1e-3


def test_stream_kernel_shards_and_status():
    K, B, N = 4096, 4, 600
    g = torch.Generator(device=DEV).manual_seed(5)
    lc, lu = torch.randn(B, N, K, device=DEV, generator=g), torch.randn(B, N, K, device=DEV, generator=g)
    t = torch.tensor([3, 50, 70, 99], device=DEV)
    x_t = torch.randint(0, K + 1, (B, N), device=DEV, generator=g)
    whole = _run(lc, lu, x_t, t, K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, seed=9, offset=5)
    parts = [_run(lc[b:e], lu[b:e], x_t[b:e].contiguous(), t[b:e].contiguous(), K, _lib.SAMPLE_PHILOX,
                  _lib.KERNEL_STREAM, seed=9, offset=5, row_offset=b * N) for b, e in ((0, 1), (1, 4))]
    assert torch.equal(torch.cat(parts), whole)
    status = ops.new_status(DEV)
    x_bad = x_t.clone()
    x_bad[2, 17] = -1
    _run(lc, lu, x_bad, t, K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, status=status)
    assert int(status.item()) & _lib.STATUS_BAD_TOKEN
    status.zero_()
    _run(lc, lu, x_t, torch.tensor([3, 50, 700, 99], device=DEV), K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, status=status)
    assert int(status.item()) & _lib.STATUS_BAD_T
    with pytest.raises(Exception):  # outputs are not something the stream kernel produces
        ops.fused_step(lc, lu, x_t, t, _table(K), guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX,
                       kernel=_lib.KERNEL_STREAM, want_post=True)


def test_stream_sampling_statistics():
    """Chi-square of 40k production draws of one row against exp(posterior)."""
    K, R = 1024, 40000
    sched = O.make_schedule(T, K)
    lc1, lu1, _, _, _ = O.synth_inputs(1, 1, K, 50, sched, seed=800, scale=3.0)
    lc, lu = lc1.expand(1, R, K).contiguous().to(DEV), lu1.expand(1, R, K).contiguous().to(DEV)
    t = torch.full((1,), 50, dtype=torch.long, device=DEV)
    for token in (K, 7):
        x_t = torch.full((1, R), token, dtype=torch.long, device=DEV)
        draws = _run(lc, lu, x_t, t, K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, seed=31, offset=token)
        post = ops.fused_step(lc[:, :1], lu[:, :1], x_t[:, :1].contiguous(), t, _table(K), guidance_scale=2.0,
                              sample_mode=_lib.SAMPLE_NONE, want_post=True)["post"][0, 0, :K + 1]
        p = post.double().exp().cpu()
        p = p / p.sum()
        counts = torch.bincount(draws.flatten().cpu(), minlength=K + 1).double()
        order = torch.argsort(p)
        be, bc, ae, ac = [], [], 0.0, 0.0
        for e_, c_ in zip((p[order] * R).numpy(), counts[order].numpy()):
            ae, ac = ae + e_, ac + c_
            if ae >= 5:
                be.append(ae), bc.append(ac)
                ae = ac = 0.0
        be[-1] += ae
        bc[-1] += ac
        chi2 = float((((np.array(bc) - np.array(be)) ** 2) / np.array(be)).sum())
        dof = len(be) - 1
        assert chi2 < dof + 5 * np.sqrt(2 * dof) + 10, (chi2, dof)


def test_philox7_uniforms_many_rows_and_offsets():
    """The noise itself (Philox4x32-7, 23-bit draws): Kolmogorov-Smirnov against U(0,1) per row and pooled, over 96
    distinct rows x 3 step offsets, moments, and the correlation between neighbouring classes, rows and offsets."""
    from scipy import stats
    K, R = 4096, 96
    us = []
    for off in (0, 1, 2**33 + 5):
        u = ops.philox_uniform(1, R, K, seed=0xD3, offset=off, row_offset=7_000_000_000, device=DEV)[0, :, :K + 1].double().cpu().numpy()
        us.append(u)
        assert u.min() > 0.0 and u.max() < 1.0
        worst_p = min(stats.kstest(u[r], "uniform").pvalue for r in range(R))
        assert worst_p > 1e-4 / R, worst_p                      # Bonferroni over the rows
        assert stats.kstest(u.ravel(), "uniform").pvalue > 1e-3  # ~4e5 draws pooled
        n = u.size
        assert abs(u.mean() - 0.5) < 5 * np.sqrt(1 / 12 / n) and abs(u.var() - 1 / 12) < 5 * np.sqrt(1 / 180 / n)
        for a_, b_ in ((u[:, :-1], u[:, 1:]), (u[:-1], u[1:]), (u[:, :-4], u[:, 4:]), (u[:, :-512], u[:, 512:])):
            r_ = np.corrcoef(a_.ravel(), b_.ravel())[0, 1]      # classes k / k+1, rows r / r+1, chunk and call partners
            assert abs(r_) < 5 / np.sqrt(a_.size), r_
    for a_, b_ in ((us[0], us[1]), (us[1], us[2])):             # consecutive steps of a chain
        assert abs(np.corrcoef(a_.ravel(), b_.ravel())[0, 1]) < 5 / np.sqrt(a_.size)
    # the 23-bit lattice: u = 1 - (2m+1)/2^24, every draw an odd multiple of 2^-24
    m = np.round((1.0 - us[0]) * 2**24).astype(np.int64)
    assert (m % 2 == 1).all()
    expect_distinct = 2**23 * (1 - np.exp(-m.size / 2**23))       # birthday bound for 23-bit draws
    assert abs(len(np.unique(m >> 1)) / expect_distinct - 1) < 0.01


def test_sampler_statistics_many_rows_and_timesteps():
    """Per-class counts of the production sampler over 64 DISTINCT rows (own logits, masked and unmasked x_t, four
    timesteps), 3000 draws each, against exp(posterior) of the oracle: one pooled chi-square plus a bound on every row's
    own statistic.  Each replica of a row sits at a different global row index, i.e. draws fresh noise."""
    K, ROWS, R = 1024, 64, 3000
    sched = O.make_schedule(T, K)
    g = torch.Generator().manual_seed(4321)
    base_c, base_u = torch.randn(4, ROWS // 4, K, generator=g) * 2.5, torch.randn(4, ROWS // 4, K, generator=g) * 2.5
    t = torch.tensor([0, 30, 60, 99])
    x_base = torch.randint(0, K, (4, ROWS // 4), generator=g)
    x_base[:, ::2] = K                                           # half the rows masked
    lc = base_c.repeat_interleave(R, dim=1).to(DEV)              # [4, 16 * R, K]: R replicas of each distinct row
    lu = base_u.repeat_interleave(R, dim=1).to(DEV)
    x_t = x_base.repeat_interleave(R, dim=1).to(DEV)
    draws = _run(lc, lu, x_t, t.to(DEV), K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, seed=77, offset=12).cpu()
    del lc, lu
    _, post, _ = O.p_sample_step(sched, base_c.permute(0, 2, 1), base_u.permute(0, 2, 1), O.index_to_log_onehot(x_base, K + 1),
                                 t, 2.0, torch.rand(4, K + 1, ROWS // 4, generator=g))
    p_all = post.double().exp().permute(0, 2, 1)                 # [4, 16, K+1]
    total_chi2, total_dof = 0.0, 0
    for b in range(4):
        for r in range(ROWS // 4):
            p = p_all[b, r] / p_all[b, r].sum()
            counts = torch.bincount(draws[b, r * R:(r + 1) * R], minlength=K + 1).double()
            order = torch.argsort(p)
            be, bc, ae, ac = [], [], 0.0, 0.0
            for e_, c_ in zip((p[order] * R).numpy(), counts[order].numpy()):
                ae, ac = ae + e_, ac + c_
                if ae >= 5:
                    be.append(ae), bc.append(ac)
                    ae = ac = 0.0
            be[-1] += ae
            bc[-1] += ac
            chi2 = float((((np.array(bc) - np.array(be)) ** 2) / np.array(be)).sum())
            dof = len(be) - 1
            assert chi2 < dof + 6 * np.sqrt(2 * dof) + 10, (b, r, chi2, dof)
            total_chi2, total_dof = total_chi2 + chi2, total_dof + dof
    z = (total_chi2 - total_dof) / np.sqrt(2 * total_dof)
    print(f"[sampler statistics] 64 rows x {R} draws: pooled chi2 {total_chi2:.0f} on {total_dof} dof (z = {z:+.2f})")
    assert abs(z) < 5


def test_tensors_on_a_non_current_device():
    """The C ABI makes the device that owns the tensors current for the call (and restores the caller's): a model on
    cuda:1 while cuda:0 is current launches on cuda:1 with cuda:1's SM count."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    K, B, N = 4096, 2, 1024
    other = torch.device("cuda", 1)
    g = torch.Generator().manual_seed(3)
    lc, lu = torch.randn(B, N, K, generator=g), torch.randn(B, N, K, generator=g)
    x_t = torch.randint(0, K + 1, (B, N), generator=g)
    t = torch.tensor([10, 90])
    sched = O.pack_schedule(O.make_schedule(T, K))
    torch.cuda.set_device(0)
    want = ops.fused_step(lc.to(DEV), lu.to(DEV), x_t.to(DEV), t.to(DEV), ops.build_coef_table(sched.to(DEV), T, K),
                          guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX, seed=1, offset=2)["x_prev"].cpu()
    got = ops.fused_step(lc.to(other), lu.to(other), x_t.to(other), t.to(other), ops.build_coef_table(sched.to(other), T, K),
                         guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX, seed=1, offset=2)["x_prev"]
    assert got.device == other and torch.cuda.current_device() == 0
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("kernel", [_lib.KERNEL_STREAM, _lib.KERNEL_ROWS])
def test_no_access_outside_the_row_windows(kernel):
    """compute-sanitizer is not available on the pool, so fence the buffers instead: NaN guard bands around
    pitched input rows (an out-of-window read would poison the softmax and change tokens) and sentinels around the
    token output (an out-of-window write would clobber them)."""
    K, B, N, PITCH, PAD = 4096, 2, 600, 4096 + 64, 1024
    g = torch.Generator(device=DEV).manual_seed(17)
    buf_c = torch.full((PAD + B * N * PITCH + PAD,), float("nan"), device=DEV)
    buf_u = torch.full_like(buf_c, float("nan"))
    lc = buf_c[PAD:PAD + B * N * PITCH].view(B, N, PITCH)[:, :, :K]
    lu = buf_u[PAD:PAD + B * N * PITCH].view(B, N, PITCH)[:, :, :K]
    lc.copy_(torch.randn(B, N, K, device=DEV, generator=g))
    lu.copy_(torch.randn(B, N, K, device=DEV, generator=g))
    x_t = torch.randint(0, K + 1, (B, N), device=DEV, generator=g)
    t = torch.tensor([10, 90], device=DEV)
    out_buf = torch.full((64 + B * N + 64,), -7, dtype=torch.int64, device=DEV)
    x_prev = out_buf[64:64 + B * N].view(B, N)
    ops.fused_step(lc, lu, x_t, t, _table(K), guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX, seed=3, offset=1,
                   kernel=kernel, x_prev_out=x_prev)
    torch.cuda.synchronize()
    assert (out_buf[:64] == -7).all() and (out_buf[-64:] == -7).all()
    dense = _run(lc.contiguous(), lu.contiguous(), x_t, t, K, _lib.SAMPLE_PHILOX, kernel, seed=3, offset=1)
    assert torch.equal(dense, x_prev)          # pitched rows == dense rows: nothing outside [0, K) was read
    assert int(x_prev.min()) >= 0 and int(x_prev.max()) <= K


def test_stream_second_attempt_and_exhaustive_fallback_at_scale():
    """Rows whose best survivor misses the acceptance bound get a second attempt (thinned at c = 16, survivors scored on
    the spot) and, should that fail too, exhaustive scoring.  thin_factor 0.5 sends most rows to the second attempt,
    1e-3 (the documented test knob) makes both attempts fail: the tokens never change.  Guidance off, K = 1024, more
    than two score batches per group."""
    K, B, N = 1024, 20, 4096
    g = torch.Generator(device=DEV).manual_seed(77)
    lc = torch.randn(B, N, K, device=DEV, generator=g)
    x_t = torch.randint(0, K + 1, (B, N), device=DEV, generator=g)
    t = torch.randint(0, T, (B,), device=DEV, generator=g)
    kw = dict(s=0.0, seed=5, offset=8)
    status = ops.new_status(DEV)
    exact = _run(lc, None, x_t, t, K, _lib.SAMPLE_PHILOX_EXACT, _lib.KERNEL_STREAM, **kw)
    for c in (1e-3, 0.5, 2.0, 0.0):
        status.zero_()
        got = _run(lc, None, x_t, t, K, _lib.SAMPLE_PHILOX, _lib.KERNEL_STREAM, thin_factor=c, status=status, **kw)
        assert torch.equal(got, exact), c
        assert int(status.item()) & _lib.STATUS_FALLBACK   # 80k rows: some are redone even at the default c


def test_steps_can_be_captured_in_a_cuda_graph():
    """The step allocates nothing and never synchronises, so a chain of steps (tokens fed back, one Philox offset per
    step) can be captured once and replayed: same tokens as the eager chain."""
    K, B, N, STEPS = 4096, 2, 1024, 4
    g = torch.Generator(device=DEV).manual_seed(9)
    lc, lu = torch.randn(B, N, K, device=DEV, generator=g), torch.randn(B, N, K, device=DEV, generator=g)
    table = _table(K)
    ts = [torch.full((B,), 99 - i, dtype=torch.long, device=DEV) for i in range(STEPS)]
    x_a = torch.full((B, N), K, dtype=torch.long, device=DEV)
    x_b = torch.empty_like(x_a)
    status = ops.new_status(DEV)

    def chain():
        src, dst = x_a, x_b
        for i in range(STEPS):
            ops.fused_step(lc, lu, src, ts[i], table, guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX, seed=3, offset=i,
                           kernel=_lib.KERNEL_STREAM, x_prev_out=dst, status=status)
            src, dst = dst, src
        return src

    eager = chain().clone()
    x_a.fill_(K)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            out = chain()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2):
        x_a.fill_(K)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,N,K,tval,s,scale", [
    (2, 1024, 4096, 50, 2.0, 1.0),
    (3, 700, 4096, [0, 42, 99], 3.0, 4.0),
    (2, 900, 4096, 30, 2.0, 12.0),            # general (clamping) path
    (2, 1100, 4096, 40, None, 1.0),           # guidance off: 64 classes per thread
    (2, 1300, 2048, 60, 2.0, 1.0),
    (2, 1500, 2048, 25, None, 1.0),
    (1, 2100, 1024, 33, 2.0, 1.0),
    (1, 37, 4096, 50, 2.0, 1.0),
])
def test_stream_kernel_16bit_logits(dtype, B, N, K, tval, s, scale):
    """A denoiser under autocast hands over float16 / bfloat16 logits.  The stream kernel reads them in place (half the
    HBM traffic) and widens in registers, so its tokens AND its posterior must be bit-identical to the fp32 kernel run on
    `logits.float()` - which is what the reference computes (`log_softmax(out.double())`, diffusion_transformer.py:231)."""
    sched = O.make_schedule(T, K)
    g = torch.Generator(device=DEV).manual_seed(B * 77 + N)
    lc16 = (torch.randn(B, N, K, device=DEV, generator=g) * scale).to(dtype)
    lu16 = None if s is None else (torch.randn(B, N, K, device=DEV, generator=g) * scale).to(dtype)
    lc32, lu32 = lc16.float(), None if lu16 is None else lu16.float()
    t = (torch.tensor(tval) if isinstance(tval, list) else torch.full((B,), tval)).long().to(DEV)
    p_mask = sched["log_cumprod_ct"][t.cpu()].exp().view(B, 1).to(DEV)
    x_t = torch.where(torch.rand(B, N, device=DEV, generator=g) < p_mask, torch.full((B, N), K, device=DEV),
                      torch.randint(0, K, (B, N), device=DEV, generator=g))
    kw = dict(guidance_scale=0.0 if s is None else s, seed=7, offset=11, row_offset=555)
    for mode, tf in ((_lib.SAMPLE_PHILOX, 0.0), (_lib.SAMPLE_PHILOX, 1.0), (_lib.SAMPLE_PHILOX, 1e-3), (_lib.SAMPLE_PHILOX_EXACT, 0.0)):
        a = ops.fused_step(lc16, lu16, x_t, t, _table(K), sample_mode=mode, thin_factor=tf, want_winner_post=True, **kw)
        b = ops.fused_step(lc32, lu32, x_t, t, _table(K), sample_mode=mode, thin_factor=tf, want_winner_post=True,
                           kernel=_lib.KERNEL_STREAM, **kw)
        assert torch.equal(a["x_prev"], b["x_prev"]), (mode, tf)
        assert torch.equal(a["winner_post"], b["winner_post"]), (mode, tf)
    # without the verification output (the production call), AUTO picks the stream kernel for 16-bit rows of any count
    plain = ops.fused_step(lc16, lu16, x_t, t, _table(K), sample_mode=_lib.SAMPLE_PHILOX, **kw)["x_prev"]
    assert torch.equal(plain, b["x_prev"])
    # the oracle on the up-cast logits, a sample of rows
    rows = torch.arange(0, N, max(1, N // 16))
    u = ops.philox_uniform(B, N, K, seed=7, offset=11, row_offset=555, device=DEV)[:, rows, :K + 1].cpu().permute(0, 2, 1)
    out_o, post_o, _ = O.p_sample_step(sched, lc32[:, rows].cpu().permute(0, 2, 1),
                                       None if lu32 is None else lu32[:, rows].cpu().permute(0, 2, 1),
                                       O.index_to_log_onehot(x_t[:, rows].cpu(), K + 1), t.cpu(), kw["guidance_scale"], u)
    H.assert_tokens_match(plain[:, rows].cpu().numpy(), out_o.argmax(1).numpy(), O.near_ties(post_o, u).numpy(), f"{dtype} stream vs oracle")
    want = post_o.gather(1, plain[:, rows].cpu().unsqueeze(1)).squeeze(1)
    assert (a["winner_post"][:, rows].cpu() - want).abs().max().item() <= H.POST_TOL
    # what the 16-bit path does not cover is refused, never silently converted
    with pytest.raises(Exception):
        ops.fused_step(lc16, lu16, x_t, t, _table(K), sample_mode=_lib.SAMPLE_NONE, want_post=True, **kw)
    with pytest.raises(Exception):
        ops.fused_step(lc16, lu16, x_t, t, _table(K), sample_mode=_lib.SAMPLE_PHILOX, kernel=_lib.KERNEL_ROWS, **kw)

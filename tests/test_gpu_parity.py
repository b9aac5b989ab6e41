"""Parity of the CUDA path (through the C ABI) against the golden vectors and the oracle.  Needs a B200."""
import hashlib

import numpy as np
import pytest
import torch

import d3pm_b200
from d3pm_b200 import _lib, ops
from oracle import d3pm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _table(T, K):
    sched = O.pack_schedule(O.make_schedule(T, K)).to(DEV)
    return ops.build_coef_table(sched, T, K)


def _uniform_rows(u_tm):
    """token-major uniform [B,N,K+1] (numpy / cpu tensor) -> padded device rows [B,N,pitch]."""
    u_tm = torch.as_tensor(u_tm)
    B, N, C = u_tm.shape
    rows = ops.alloc_rows(B, N, C, DEV)
    rows.fill_(0.5)
    rows[:, :, :C] = u_tm.to(DEV)
    return rows


def _cuda_step(lc, lu, x_t, t, T, u_tm=None, s=2.0, mode=_lib.SAMPLE_GUMBEL, **kw):
    K = lc.shape[-1]
    out = ops.fused_step(
        torch.as_tensor(lc).to(DEV), None if lu is None else torch.as_tensor(lu).to(DEV),
        torch.as_tensor(x_t).to(DEV), torch.as_tensor(t).to(DEV), _table(T, K), guidance_scale=s, sample_mode=mode,
        gumbel=None if u_tm is None else _uniform_rows(u_tm), gumbel_is_uniform=True, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items()}


# ------------------------------------------------------------------ golden vectors (reference outputs)
@pytest.mark.parametrize("path", H.step_fixtures(), ids=lambda p: p.split("step_")[-1][:-4])
def test_fused_step_against_golden(path):
    fx = H.load(path)
    K, T = int(fx["K"]), int(fx["T"])
    s = H.fixture_guidance(fx)
    out = _cuda_step(fx["logits_c"], None if s is None else fx["logits_u"], fx["x_t"], fx["t"], T, fx["uniform"],
                     s=0.0 if s is None else s, want_post=True, want_recon=True, want_gap=True)
    post, recon = out["post"][:, :, :K + 1].numpy(), out["recon"][:, :, :K + 1].numpy()
    assert np.abs(post - fx["post"]).max() <= H.POST_TOL
    assert np.abs(recon - fx["recon"]).max() <= H.POST_TOL
    H.assert_tokens_match(out["x_prev"].numpy(), fx["x_prev"], fx["near_tie"], path)
    # the kernel's own near-tie log agrees with the reference-side one
    assert not ((out["gap"].numpy() < H.NEAR_TIE_GAP / 2) & ~fx["near_tie"]).any()


def test_config1_full_size_against_reference_digest():
    """BASELINE config 1: B=1, 16x8x8 grid, K=4096, s=2, t=50 — against the reference's own outputs."""
    fx = H.load(f"{H.GOLDEN}/step_config1_digest.npz")
    T, K = int(fx["T"]), int(fx["K"])
    sched = O.make_schedule(T, K)
    lc, lu, x_t, t, u = O.synth_inputs(1, 1024, K, 50, sched, seed=int(fx["seed"]))
    h = hashlib.sha256()
    for a in (lc, lu, x_t, t, u):
        h.update(np.ascontiguousarray(a.numpy()).tobytes())
    assert h.hexdigest() == str(fx["inputs_sha256"])
    out = _cuda_step(lc, lu, x_t, t, T, u.permute(0, 2, 1), s=2.0, want_post=True)
    H.assert_tokens_match(out["x_prev"].numpy(), fx["x_prev"], fx["near_tie"], "config1")
    post = out["post"][:, :, :K + 1]
    assert np.abs(post[0, fx["sample_rows"]].numpy() - fx["post_rows"]).max() <= H.POST_TOL
    assert np.abs(torch.logsumexp(post.double(), -1).numpy() - fx["post_row_lse"]).max() <= H.POST_TOL


def test_qposterior_onehot_golden():
    fx = H.load(f"{H.GOLDEN}/qpost_onehot_k64.npz")
    T, K = int(fx["T"]), int(fx["K"])
    x0 = torch.from_numpy(fx["x0"]).to(DEV)
    rows = ops.tokens_to_log_onehot_rows(x0, K + 1)
    post = ops.q_posterior_rows(rows, rows.shape[2], torch.from_numpy(fx["x_t"]).to(DEV),
                                torch.from_numpy(fx["t"]).to(DEV), _table(T, K), K)
    assert np.abs(post[:, :, :K + 1].cpu().numpy() - fx["post"]).max() <= H.POST_TOL


# ------------------------------------------------------------------ oracle on fresh seeded inputs
CASES = [
    # B, N, K, t, scale, spikes, guidance, seed
    (2, 16, 4096, 50, 1.0, False, 2.0, 300),
    (2, 16, 4096, 0, 1.0, False, 2.0, 310),
    (2, 16, 4096, 99, 1.0, False, 2.0, 320),
    (2, 16, 4096, 1, 8.0, True, 2.0, 330),
    (3, 8, 4096, [0, 57, 99], 8.0, True, 5.0, 340),
    (2, 8, 4096, 25, 1.0, False, None, 350),
    (2, 8, 2048, 75, 1.0, False, 2.0, 360),
    (2, 8, 8192, 40, 1.0, False, 2.0, 370),
    (2, 8, 1000, 10, 2.0, True, 1.5, 380),   # K not a multiple of the CTA width
    (1, 1, 8, 3, 1.0, False, 2.0, 390),      # smallest legal row
    (2, 5, 4096, 50, 30.0, False, 2.0, 400),  # large-magnitude logits
]


@pytest.mark.parametrize("B,N,K,tval,scale,spikes,s,seed", CASES)
def test_fused_step_against_oracle(B, N, K, tval, scale, spikes, s, seed):
    T = 100
    sched = O.make_schedule(T, K)
    t_in = torch.tensor(tval) if isinstance(tval, list) else tval
    lc, lu, x_t, t, u = O.synth_inputs(B, N, K, t_in, sched, seed=seed, scale=scale, spikes=spikes)
    out_o, post_o, recon_o = O.p_sample_step(sched, lc.permute(0, 2, 1), None if s is None else lu.permute(0, 2, 1),
                                             O.index_to_log_onehot(x_t, K + 1), t, 0.0 if s is None else s, u)
    out = _cuda_step(lc, None if s is None else lu, x_t, t, T, u.permute(0, 2, 1), s=0.0 if s is None else s,
                     want_post=True, want_recon=True)
    assert (out["post"][:, :, :K + 1] - post_o.permute(0, 2, 1)).abs().max() <= H.POST_TOL
    assert (out["recon"][:, :, :K + 1] - recon_o.permute(0, 2, 1)).abs().max() <= H.POST_TOL
    H.assert_tokens_match(out["x_prev"].numpy(), out_o.argmax(1).numpy(), O.near_ties(post_o, u).numpy(), f"seed{seed}")


def test_closed_form_edge_cases():
    """t = 0 wraps to the identity slot: a masked token's posterior is p(x0) itself and [MASK] gets log 1e-30;
    posterior rows are normalised wherever no clamp fired (SURVEY §8 c)."""
    T, K, B, N = 100, 4096, 2, 32
    sched = O.make_schedule(T, K)
    lc, lu, _, _, _ = O.synth_inputs(B, N, K, 0, sched, seed=500)
    x_t = torch.full((B, N), K, dtype=torch.long)
    x_t[:, ::2] = torch.randint(0, K, (B, N // 2), generator=torch.Generator().manual_seed(1))
    t = torch.zeros(B, dtype=torch.long)
    out = _cuda_step(lc, lu, x_t, t, T, None, mode=_lib.SAMPLE_NONE, want_post=True, want_recon=True)
    post, recon = out["post"][:, :, :K + 1], out["recon"][:, :, :K + 1]
    masked = x_t == K
    assert (post[masked][:, :K] - recon[masked][:, :K]).abs().max() <= 2e-5
    assert (post[masked][:, K] - O.LOG_TINY).abs().max() <= 1e-4
    assert (recon[:, :, K] == -70).all()
    lse = torch.logsumexp(post.double(), -1)
    assert lse.abs().max() <= 1e-4
    # an unmasked token at a middle step keeps almost all of its mass
    t50 = torch.full((B,), 50, dtype=torch.long)
    post50 = _cuda_step(lc, lu, x_t, t50, T, None, mode=_lib.SAMPLE_NONE, want_post=True)["post"]
    keep = post50[~masked].gather(1, x_t[~masked].unsqueeze(1)).exp()
    assert keep.min() > 0.9


# ------------------------------------------------------------------ in-kernel Philox sampling
@pytest.mark.parametrize("B,N,K,tval,s", [(2, 64, 4096, 50, 2.0), (2, 64, 4096, 0, 2.0), (2, 64, 4096, 99, 2.0),
                                          (3, 32, 4096, [1, 30, 80], 3.0), (2, 32, 2048, 50, None), (2, 16, 64, 20, 2.0)])
def test_philox_thinned_equals_exact_equals_reference_formula(B, N, K, tval, s):
    """Production sampling (thinned exponential race) == exhaustive log-space scoring == the oracle fed the
    very uniforms the kernel drew (dumped by d3pm_philox_uniform) through the reference formula."""
    T, seed, offset, row_offset = 100, 1234567, 42, 1000
    sched = O.make_schedule(T, K)
    t_in = torch.tensor(tval) if isinstance(tval, list) else tval
    lc, lu, x_t, t, _ = O.synth_inputs(B, N, K, t_in, sched, seed=600)
    kw = dict(s=0.0 if s is None else s, seed=seed, offset=offset, row_offset=row_offset)
    lu_in = None if s is None else lu
    status = ops.new_status(DEV)
    thin = _cuda_step(lc, lu_in, x_t, t, T, mode=_lib.SAMPLE_PHILOX, status=status, **kw)["x_prev"]
    exact = _cuda_step(lc, lu_in, x_t, t, T, mode=_lib.SAMPLE_PHILOX_EXACT, want_post=True, want_gap=True, **kw)
    assert torch.equal(thin, exact["x_prev"])
    forced = _cuda_step(lc, lu_in, x_t, t, T, mode=_lib.SAMPLE_PHILOX, thin_factor=1e-3, status=status, **kw)["x_prev"]
    assert torch.equal(forced, thin)                    # the exhaustive fallback gives the same draw
    assert int(status.item()) & _lib.STATUS_FALLBACK
    u = ops.philox_uniform(B, N, K, seed=seed, offset=offset, row_offset=row_offset, device=DEV)[:, :, :K + 1].cpu()
    assert u.min() > 0 and u.max() < 1
    out_o, post_o, _ = O.p_sample_step(sched, lc.permute(0, 2, 1), None if s is None else lu.permute(0, 2, 1),
                                       O.index_to_log_onehot(x_t, K + 1), t, 0.0 if s is None else s, u.permute(0, 2, 1))
    H.assert_tokens_match(thin.numpy(), out_o.argmax(1).numpy(), O.near_ties(post_o, u.permute(0, 2, 1)).numpy(), "philox")
    # different offset -> different noise; different shard offset -> different rows
    other = _cuda_step(lc, lu_in, x_t, t, T, mode=_lib.SAMPLE_PHILOX, **{**kw, "offset": offset + 1})["x_prev"]
    if not bool((thin == K).all()):  # (at t = 99 nearly every token stays [MASK] whatever the noise)
        assert not torch.equal(other, thin)
    u2 = ops.philox_uniform(B, N, K, seed=seed, offset=offset + 1, row_offset=row_offset, device=DEV)[:, :, :K + 1].cpu()
    assert (u2 != u).float().mean() > 0.99


def test_sharded_rows_reproduce_the_single_gpu_stream():
    T, K, B, N = 100, 4096, 4, 32
    sched = O.make_schedule(T, K)
    lc, lu, x_t, t, _ = O.synth_inputs(B, N, K, torch.tensor([3, 50, 70, 99]), sched, seed=700)
    whole = _cuda_step(lc, lu, x_t, t, T, mode=_lib.SAMPLE_PHILOX, seed=9, offset=5)["x_prev"]
    parts = [_cuda_step(lc[b:e], lu[b:e], x_t[b:e], t[b:e], T, mode=_lib.SAMPLE_PHILOX, seed=9, offset=5,
                        row_offset=b * N)["x_prev"] for b, e in ((0, 1), (1, 4))]
    assert torch.equal(torch.cat(parts), whole)


def test_philox_sampling_follows_the_posterior():
    """Chi-square of 20k draws of one row against exp(posterior) (classes pooled to >= 5 expected)."""
    T, K, R = 100, 64, 20000
    sched = O.make_schedule(T, K)
    lc1, lu1, _, _, _ = O.synth_inputs(1, 1, K, 50, sched, seed=800, scale=2.0)
    lc, lu = lc1.expand(1, R, K).contiguous(), lu1.expand(1, R, K).contiguous()
    for token in (K, 7):
        x_t = torch.full((1, R), token, dtype=torch.long)
        t = torch.full((1,), 50, dtype=torch.long)
        out = _cuda_step(lc, lu, x_t, t, T, mode=_lib.SAMPLE_PHILOX, seed=77, offset=token)
        post = _cuda_step(lc[:, :1], lu[:, :1], x_t[:, :1], t, T, mode=_lib.SAMPLE_NONE, want_post=True)["post"][0, 0, :K + 1]
        p = post.double().exp()
        p = p / p.sum()
        counts = torch.bincount(out["x_prev"].flatten(), minlength=K + 1).double()
        order = torch.argsort(p)
        exp_sorted, cnt_sorted = (p[order] * R).numpy(), counts[order].numpy()
        bins_e, bins_c, acc_e, acc_c = [], [], 0.0, 0.0
        for e_, c_ in zip(exp_sorted, cnt_sorted):
            acc_e, acc_c = acc_e + e_, acc_c + c_
            if acc_e >= 5:
                bins_e.append(acc_e), bins_c.append(acc_c)
                acc_e = acc_c = 0.0
        bins_e[-1] += acc_e
        bins_c[-1] += acc_c
        chi2 = float((((np.array(bins_c) - np.array(bins_e)) ** 2) / np.array(bins_e)).sum())
        dof = len(bins_e) - 1
        assert chi2 < dof + 5 * np.sqrt(2 * dof) + 10, (chi2, dof)


# ------------------------------------------------------------------ status word / error behaviour
def test_out_of_range_inputs_set_the_status_word():
    T, K = 100, 64
    sched = O.make_schedule(T, K)
    lc, lu, x_t, t, _ = O.synth_inputs(2, 4, K, 50, sched, seed=900)
    status = ops.new_status(DEV)
    x_bad = x_t.clone()
    x_bad[0, 0] = K + 5
    _cuda_step(lc, lu, x_bad, t, T, mode=_lib.SAMPLE_PHILOX, status=status)
    assert int(status.item()) & _lib.STATUS_BAD_TOKEN
    status.zero_()
    _cuda_step(lc, lu, x_t, torch.tensor([100, 5]), T, mode=_lib.SAMPLE_PHILOX, status=status)
    assert int(status.item()) & _lib.STATUS_BAD_T
    with pytest.raises(d3pm_b200.D3PMError):
        ops.fused_step(lc.to(DEV)[:, :, :6], None, x_t.to(DEV), t.to(DEV), _table(T, K), guidance_scale=2.0,
                       sample_mode=_lib.SAMPLE_PHILOX)  # K % 4 != 0 is refused, not mis-computed


# ------------------------------------------------------------------ fine-grained operators
def test_index_log_onehot_round_trip_both_layouts():
    K, B, N = 4096, 2, 40
    x = torch.randint(0, K + 1, (B, N), generator=torch.Generator().manual_seed(3)).to(DEV)
    rows = ops.tokens_to_log_onehot_rows(x, K + 1)
    logical = ops.as_logical(rows, K + 1)
    ref = O.index_to_log_onehot(x.cpu(), K + 1)
    assert torch.equal(logical.cpu(), ref)
    assert torch.equal(ops.argmax_classes(logical), x)                       # token-major view
    assert torch.equal(ops.argmax_classes(logical.contiguous()), x)          # reference-contiguous layout
    rows2, pitch = ops.to_rows(logical.contiguous())
    assert pitch == 4100 and torch.equal(rows2, rows[:, :, :K + 1])


def test_gumbel_argmax_operator():
    K, B, N = 4096, 2, 24
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(B, K + 1, N, generator=g) * 3).clamp(-70, 0)
    u = torch.rand(B, K + 1, N, generator=g)
    want = O.log_sample_categorical(logits, u, return_index=True)
    rows, pitch = ops.to_rows(logits.to(DEV))
    urows, upitch = ops.to_rows(u.to(DEV))
    got, gap = ops.gumbel_argmax_rows(rows, pitch, K + 1, noise_rows=urows, pitch_noise=upitch, noise_kind=1, want_gap=True)
    H.assert_tokens_match(got.cpu().numpy(), want.numpy(), O.near_ties(logits, u).numpy(), "gumbel_argmax")
    top2 = (O.gumbel_from_uniform(u) + logits).topk(2, dim=1).values
    assert (gap.cpu() - (top2[:, 0] - top2[:, 1])).abs().max() < 1e-4
    # Philox noise: same stream as the fused step's dump
    got_p = ops.gumbel_argmax_rows(rows, pitch, K + 1, noise_kind=2, seed=11, offset=3)
    up = ops.philox_uniform(B, N, K, seed=11, offset=3, device=DEV)[:, :, :K + 1].cpu().permute(0, 2, 1)
    want_p = O.log_sample_categorical(logits, up, return_index=True)
    H.assert_tokens_match(got_p.cpu().numpy(), want_p.numpy(), O.near_ties(logits, up).numpy(), "gumbel_argmax philox")


# ------------------------------------------------------------------ full-size properties (BASELINE config 2 shape)
def test_full_size_properties_config2():
    """B=16, 16x16x16 grid, K=4096, s=2, t=50: too big for the CPU oracle in seconds, so check
    size-independent properties: production sampling == exhaustive scoring; posterior rows normalised;
    a 64-row sample agrees with the oracle; nothing stays out of range."""
    T, K, B, N = 100, 4096, 16, 4096
    g = torch.Generator(device=DEV).manual_seed(0)
    lc = torch.randn(B, N, K, device=DEV, generator=g)
    lu = torch.randn(B, N, K, device=DEV, generator=g)
    sched = O.make_schedule(T, K)
    t = torch.full((B,), 50, dtype=torch.long, device=DEV)
    x_t = torch.where(torch.rand(B, N, device=DEV, generator=g) < float(sched["log_cumprod_ct"][50].exp()),
                      torch.full((B, N), K, device=DEV), torch.randint(0, K, (B, N), device=DEV, generator=g))
    table = _table(T, K)
    kw = dict(guidance_scale=2.0, seed=2024, offset=7)
    thin = ops.fused_step(lc, lu, x_t, t, table, sample_mode=_lib.SAMPLE_PHILOX, **kw)["x_prev"]
    exact = ops.fused_step(lc, lu, x_t, t, table, sample_mode=_lib.SAMPLE_PHILOX_EXACT, **kw)["x_prev"]
    assert torch.equal(thin, exact)
    assert int(thin.min()) >= 0 and int(thin.max()) <= K
    sel = slice(0, 4)
    post = ops.fused_step(lc[:1, sel], lu[:1, sel], x_t[:1, sel].contiguous(), t[:1], table, sample_mode=_lib.SAMPLE_NONE,
                          want_post=True, guidance_scale=2.0)["post"][:, :, :K + 1]
    assert torch.logsumexp(post.double(), -1).abs().max() < 1e-4
    # 2048 rows against the oracle (every 4th row of the first two videos): tokens with the dumped uniforms, and the
    # posterior log-prob of the sampled class AS THE PRODUCTION (stream) KERNEL COMPUTED IT (winner_post) against the
    # oracle's posterior row: the 1e-4 gate measured on the kernel that is benchmarked
    wp = ops.fused_step(lc, lu, x_t, t, table, sample_mode=_lib.SAMPLE_PHILOX, want_winner_post=True, **kw)
    assert torch.equal(wp["x_prev"], thin)
    rows = torch.arange(0, N, 4)
    u = ops.philox_uniform(B, N, K, seed=2024, offset=7, device=DEV)
    lc_s, lu_s = lc[:2, rows].cpu(), lu[:2, rows].cpu()
    x_s, u_s = x_t[:2, rows].cpu(), u[:2, rows, :K + 1].cpu().permute(0, 2, 1)
    del u
    out_o, post_o, _ = O.p_sample_step(sched, lc_s.permute(0, 2, 1), lu_s.permute(0, 2, 1),
                                       O.index_to_log_onehot(x_s, K + 1), t[:2].cpu(), 2.0, u_s)
    tok_o = out_o.argmax(1)
    assert tok_o.numel() >= 2048
    H.assert_tokens_match(thin[:2, rows].cpu().numpy(), tok_o.numpy(), O.near_ties(post_o, u_s).numpy(), "config2")
    got_tok = thin[:2, rows].cpu()
    want_lp = post_o.gather(1, got_tok.unsqueeze(1)).squeeze(1)  # the oracle's posterior at the class the kernel drew
    err = (wp["winner_post"][:2, rows].cpu() - want_lp).abs().max().item()
    print(f"[config 2] stream kernel: posterior of the sampled class vs oracle over {tok_o.numel()} rows: max |err| = {err:.2e}")
    assert err <= H.POST_TOL


@pytest.mark.parametrize("C,N", [(4097, 700), (2049, 300), (65, 515), (8193, 40)])
def test_thinned_gumbel_argmax_equals_the_exhaustive_one(C, N):
    """log_sample_categorical with the library's own noise: the thinned race (softmax statistics, coarse noise filter,
    exact scoring of the survivors) draws exactly the class the per-class kernel draws - which `want_gap` still selects -
    for peaked, flat, clamped (-70) and partly -inf rows, a row of all -inf, and the reference's pitch C (no padding)."""
    B = 2
    g = torch.Generator(device=DEV).manual_seed(C + N)
    rows = ops.alloc_rows(B, N, C, DEV)
    lg = torch.randn(B, N, C, device=DEV, generator=g) * 3.0
    lg = torch.log_softmax(lg, dim=-1).clamp(-70.0, 0.0)
    lg[0, ::5] *= 0.01                                   # nearly flat rows
    lg[1, ::7, : C // 2] = float("-inf")                 # half the classes impossible
    lg[1, 3] = float("-inf")                             # nothing possible: falls back to scoring everything
    lg[0, 11] = -70.0                                    # all clamped
    rows[:, :, :C] = lg
    kw = dict(noise_kind=2, seed=21, offset=4, row_offset=999)
    thin = ops.gumbel_argmax_rows(rows, rows.shape[2], C, **kw)
    full, _ = ops.gumbel_argmax_rows(rows, rows.shape[2], C, want_gap=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(thin, full)
    dense = lg.contiguous()                              # pitch == C (4097 floats: unaligned rows)
    assert torch.equal(ops.gumbel_argmax_rows(dense, C, C, **kw), full)
    assert int(thin.min()) >= 0 and int(thin.max()) < C

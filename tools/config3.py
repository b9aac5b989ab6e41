"""BASELINE config 3: the reverse-sampling loop with the reference's REAL denoiser.

    python tools/config3.py [--videos 16] [--grid 16 16 16] [--steps 100] [--window 0] [--json out.json]

Builds the unmodified reference `DiffusionTransformer(Text2ImageTransformer(n_layer 19, n_embd 64, n_head 16))`
(`baseline/reference_loader.py`: `/root/reference` in the build container, the staged `baseline/_ref` on the GPU box),
copies its `state_dict()` strictly into `FusedDiffusionTransformer` (same denoiser class, same weights) and times on ONE GPU

  * the reference's own `sample()` (eager PyTorch on the GPU: 146 full-tensor kernels per update, :613-626),
  * `FusedDiffusionTransformer.sample()` (denoiser unchanged, update = one `d3pm_fused_step` launch),
  * the same with `enable_fused_head()` (the `to_logits` head folded into `d3pm_head_step`, logits never stored),

plus the split of a step into denoiser time (two forwards: conditional + unconditional) and update time.
`--window W` (> 0) times only the first W reverse steps of the chain (t = T-1 ... T-W) of each arm instead of all T and
scales to T steps: what `bench.py` puts in `next_rows.config3` so that its run stays within minutes.

The attention of the reference denoiser materialises a `[B, 16, N, N]` fp32 matrix per layer (17 GB at B = 16,
N = 4096; transformer_utils.py:52-58), so at config-3 size the chain is >95 % denoiser on either arm: the update is the
part this repo replaces, and the table says how much of the chain it was and is.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_models(B, N, K, T, spatial, dev, guidance=2.0, seed=0, logit_gain=1.0):
    """-> (reference model, drop-in model) on `dev`, same weights (strict state_dict copy), eval mode."""
    import torch

    import d3pm_b200
    from baseline import reference_loader as RL

    torch.manual_seed(seed)
    den_ref = RL.build_denoiser(K, N, spatial, diffusion_step=T)
    if logit_gain != 1.0:  # random init gives |logit| ~ 0.1; a gain makes the distributions as peaked as a trained model's
        with torch.no_grad():
            den_ref.to_logits[-1].weight.mul_(logit_gain)
    ref = RL.build_reference_model(den_ref, diffusion_step=T, guidance_scale=guidance, content_seq_len=N).to(dev).eval()
    den_ours = RL.build_denoiser(K, N, spatial, diffusion_step=T)
    ours = d3pm_b200.FusedDiffusionTransformer(transformer=den_ours, diffusion_step=T, alpha_init_type="alpha1",
                                               guidance_scale=guidance, content_seq_len=N)
    missing = ours.load_state_dict(ref.state_dict(), strict=True)  # the checkpoint contract: same keys, same shapes
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref, ours.to(dev).eval()


def _sync_time(fn, dev):
    import torch
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize(dev)
    return time.perf_counter() - t0, out


def run(B=16, grid=(16, 16, 16), K=4096, T=100, window=0, dev=None, guidance=2.0, quiet=False, shipped_conditioning=True):
    """Times the three arms; returns a dict (see module docstring)."""
    import torch

    dev = dev or torch.device("cuda", 0)
    N = grid[0] * grid[1] * grid[2]
    side = 1
    while side * side < N:
        side *= 2
    ref, ours = build_models(B, N, K, T, [side, side], dev, guidance=guidance)
    g = torch.Generator(device=dev).manual_seed(11)
    cond = torch.randn(B, 77, 512, device=dev, generator=g)
    cf = torch.zeros(B, 77, 512, device=dev)  # the caller zeroes the unconditional embedding (networks/discrete_diffusion.py:49)
    text = ["a video"] * B
    steps = window if window > 0 else T
    ts = list(range(T - 1, T - 1 - steps, -1))

    # ---- reference arm: its own sample() (:568-644), or its p_sample loop (:621-626) over a window of steps
    def ref_chain(ts_, whole, cond=cond, cf=cf):
        with torch.no_grad():
            if whole:
                return ref.sample(text, None, cond, cf, filter_ratio=0)["content_token"]
            log_z = torch.log(torch.cat((torch.zeros(B, K, N, device=dev), torch.ones(B, 1, N, device=dev)), dim=1))
            for ti in ts_:
                t = torch.full((B,), ti, device=dev, dtype=torch.long)
                log_z, _ = ref.p_sample(log_z, cond, cf, t, [0] * B, ref.n_sample[ti])
            return log_z.argmax(1)

    def our_chain(ts_, whole, cond=cond, cf=cf):
        if whole:
            return ours.sample(text, None, cond, cf, filter_ratio=0)["content_token"]
        x = torch.full((B, N), K, dtype=torch.int64, device=dev)
        for ti in ts_:
            t = torch.full((B,), ti, device=dev, dtype=torch.long)
            x = ours.p_sample_tokens(x, cond, cf, t)
        return x

    # denoiser alone (two forwards per step, the part both arms share)
    x_mask = torch.full((B, N), K, dtype=torch.int64, device=dev)
    t_top = torch.full((B,), T - 1, device=dev, dtype=torch.long)

    def two_forwards():
        with torch.no_grad():
            ours.transformer(x_mask, cond, t_top)
            ours.transformer(x_mask, cf, t_top)

    two_forwards()  # warm-up (allocator, cuBLAS handles)
    den_s = min(_sync_time(two_forwards, dev)[0] for _ in range(2))

    whole = window <= 0
    ref_chain(ts[:1], False), our_chain(ts[:1], False)  # warm-up of each arm on one step
    ref_s, ref_tok = _sync_time(lambda: ref_chain(ts, whole), dev)
    ours.manual_seed(1)
    our_s, our_tok = _sync_time(lambda: our_chain(ts, whole), dev)
    ours.enable_fused_head()
    fused_ok = bool(ours.fused_head_active)
    our_chain(ts[:1], False)
    ours.manual_seed(1)
    fus_s, fus_tok = _sync_time(lambda: our_chain(ts, whole), dev)
    ours.enable_fused_head(False)
    ours.check_status()
    # ---- the conditioning the reference's pipeline actually ships: its caller zeroes BOTH text embeddings
    #      (networks/discrete_diffusion.py:25, :49), so the two denoiser passes of a step compute the same logits.  The
    #      reference runs both; the drop-in notices the bitwise-equal embeddings once per chain and runs one pass
    shipped = None
    if shipped_conditioning:
        z_c, z_u = torch.zeros(B, 1, 512, device=dev), torch.zeros(B, 1, 512, device=dev)
        ref_chain(ts[:1], False, z_c, z_u), our_chain(ts[:1], False, z_c, z_u)
        sref_s, _ = _sync_time(lambda: ref_chain(ts, whole, z_c, z_u), dev)
        ours.manual_seed(1)
        sour_s, _ = _sync_time(lambda: our_chain(ts, whole, z_c, z_u), dev)
        ours.check_status()
        shipped = {"what": "both text embeddings zeroed as networks/discrete_diffusion.py:25,49 does: the reference runs two identical "
                           "denoiser passes per step, the drop-in one (bit-identical logits for both guidance branches)",
                   "reference_ms_per_step": sref_s / steps * 1e3, "ours_ms_per_step": sour_s / steps * 1e3,
                   "reference_chain_s": sref_s / steps * T, "ours_chain_s": sour_s / steps * T, "sample_speedup": sref_s / sour_s}
    if whole:
        for tok in (ref_tok, our_tok, fus_tok):
            assert tok.shape == (B, N) and int(tok.max()) < K, "a finished chain holds no [MASK]"

    # ---- the update alone, timed directly on cached logits of one step (CUDA events): the reference's p_sample with its
    #      denoiser replaced by a module that hands back the cached logits the way Text2ImageTransformer does (a
    #      [B, K, N] view of [B, N, K], transformer_utils.py:442-443), against one d3pm_fused_step launch
    with torch.no_grad():
        lc = ours.transformer(x_mask, cond, t_top).permute(0, 2, 1)
        lu = ours.transformer(x_mask, cf, t_top).permute(0, 2, 1)

    class Cached(torch.nn.Module):
        def forward(self, x_t, cond_emb, t):
            return (lc if cond_emb is cond else lu).permute(0, 2, 1)

    real_ref, real_ours = ref.transformer, ours.transformer
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def event_ms(fn, n):
        fn()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    log_mask = torch.log(torch.cat((torch.zeros(B, K, N, device=dev), torch.ones(B, 1, N, device=dev)), dim=1))
    ref.transformer = Cached()
    with torch.no_grad():
        upd_ref_direct = event_ms(lambda: ref.p_sample(log_mask, cond, cf, t_top, [0] * B, ref.n_sample[T - 1]), 3)
    ref.transformer = real_ref
    ours.transformer = Cached()
    x_out = torch.empty_like(x_mask)
    upd_our_direct = event_ms(lambda: ours.p_sample_tokens(x_mask, cond, cf, t_top, x_prev_out=x_out), 20)
    ours.transformer = real_ours
    del log_mask

    per = lambda s: s / steps  # noqa: E731
    upd_ref, upd_our = per(ref_s) - den_s, per(our_s) - den_s
    res = {
        "what": "config 3: reverse chain with the reference's Text2ImageTransformer (n_layer 19, n_embd 64, n_head 16), "
                f"{B} videos x {N} tokens x {K}+1 classes, guidance {guidance:g}, fp32, one B200",
        "steps_timed": steps, "chain_steps": T, "window": bool(window > 0),
        "denoiser_two_forwards_ms": den_s * 1e3,
        "reference_ms_per_step": per(ref_s) * 1e3, "ours_ms_per_step": per(our_s) * 1e3,
        "ours_fused_head_ms_per_step": per(fus_s) * 1e3, "fused_head_valid": fused_ok,
        "reference_update_ms_per_step_by_subtraction": upd_ref * 1e3, "ours_update_ms_per_step_by_subtraction": upd_our * 1e3,
        "reference_chain_s": per(ref_s) * T, "ours_chain_s": per(our_s) * T, "ours_fused_head_chain_s": per(fus_s) * T,
        "sample_speedup": ref_s / our_s, "sample_speedup_fused_head": ref_s / fus_s,
        "update_speedup": upd_ref_direct / upd_our_direct,
        "reference_update_ms_direct": upd_ref_direct, "ours_update_ms_direct": upd_our_direct,
        "update_share_of_reference_step": upd_ref_direct * 1e-3 / per(ref_s), "update_share_of_our_step": upd_our_direct * 1e-3 / per(our_s),
        "token_updates_per_s_reference": B * N / per(ref_s), "token_updates_per_s_ours": B * N / per(our_s),
        "peak_memory_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
        "shipped_conditioning": shipped,
    }
    if not quiet:
        print(json.dumps(res, indent=1))
    del ref, ours
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=16)
    ap.add_argument("--grid", type=int, nargs=3, default=[16, 16, 16])
    ap.add_argument("--codes", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    out = run(a.videos, tuple(a.grid), a.codes, a.steps, a.window)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)

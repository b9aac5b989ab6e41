"""Minimal driver for ncu: a few production-mode fused steps at the BASELINE config-2 shape.

    python tools/prof_step.py [--launches 3] [--videos 16] [--mode philox|exact|gumbel|post]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import d3pm_b200  # noqa: E402
from d3pm_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--videos", type=int, default=16)
ap.add_argument("--t", type=int, default=50)
ap.add_argument("--mode", default="philox")
ap.add_argument("--no-guidance", action="store_true")
ap.add_argument("--codes", type=int, default=4096)
ap.add_argument("--thin", type=float, default=0.0, help="thinning constant c (0 = the kernel's default)")
ap.add_argument("--sleep-ms", type=float, default=0.0, help="idle gap before every launch (isolated launches)")
a = ap.parse_args()

dev = torch.device("cuda", 0)
T, K, N, B = 100, a.codes, 4096, a.videos


class _Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.content_emb = type("E", (), {"num_embed": K + 1})()


m = d3pm_b200.FusedDiffusionTransformer(transformer=_Stub(), diffusion_step=T, alpha_init_type="alpha1",
                                        guidance_scale=2.0, content_seq_len=N).to(dev)
table = m.coef_table()
g = torch.Generator(device=dev).manual_seed(0)
lc = torch.randn(B, N, K, device=dev, generator=g)
lu = None if a.no_guidance else torch.randn(B, N, K, device=dev, generator=g)
pm = float(m.log_cumprod_ct[a.t].exp())
x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < pm, torch.full((B, N), K, device=dev),
                  torch.randint(0, K, (B, N), device=dev, generator=g))
t = torch.full((B,), a.t, dtype=torch.int64, device=dev)
xp = torch.empty_like(x_t)
kw = dict(guidance_scale=2.0, seed=1, x_prev_out=xp, thin_factor=a.thin)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.launches + 1)]
ev[0].record()
for i in range(a.launches):
    if a.sleep_ms > 0:
        torch.cuda.synchronize()
        import time
        time.sleep(a.sleep_ms / 1e3)
        ev[i].record()
    if a.mode == "philox":
        ops.fused_step(lc, lu, x_t, t, table, sample_mode=_lib.SAMPLE_PHILOX, offset=i, **kw)
    elif a.mode == "exact":
        ops.fused_step(lc, lu, x_t, t, table, sample_mode=_lib.SAMPLE_PHILOX_EXACT, offset=i, **kw)
    elif a.mode == "post":
        ops.fused_step(lc, lu, x_t, t, table, sample_mode=_lib.SAMPLE_NONE, want_post=True, guidance_scale=2.0)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.launches)]
print("ms per launch:", ["%.3f" % x for x in ms], "GB/s:", "%.0f" % (B * N * ((8 * K + 16) if lu is not None else (4 * K + 16)) / min(ms) / 1e6))

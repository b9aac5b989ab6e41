"""Minimal driver for ncu / timing: a few fused head + reverse steps (d3pm_head_step) at the BASELINE config-2 shape.

    python tools/prof_head.py [--launches 3] [--videos 16] [--no-guidance]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import d3pm_b200  # noqa: E402
from d3pm_b200 import _lib, head  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--videos", type=int, default=16)
ap.add_argument("--t", type=int, default=50)
ap.add_argument("--no-guidance", action="store_true")
ap.add_argument("--stats3x", action="store_true", help="statistics pass in 3xTF32 (the round-1 kernel) instead of 1xTF32")
a = ap.parse_args()
dev = torch.device("cuda", 0)
T, K, N, B, D = 100, 4096, 4096, a.videos, 64


class _Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.content_emb = type("E", (), {"num_embed": K + 1})()


m = d3pm_b200.FusedDiffusionTransformer(transformer=_Stub(), diffusion_step=T, alpha_init_type="alpha1", guidance_scale=2.0,
                                        content_seq_len=N).to(dev)
table = m.coef_table()
g = torch.Generator(device=dev).manual_seed(0)
tl = torch.nn.Sequential(torch.nn.LayerNorm(D), torch.nn.Linear(D, K)).to(dev)
hw = head.HeadWeights.from_module(tl)
assert hw.valid
hc = torch.randn(B, N, D, device=dev, generator=g)
hu = None if a.no_guidance else torch.randn(B, N, D, device=dev, generator=g)
pm = float(m.log_cumprod_ct[a.t].exp())
x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < pm, torch.full((B, N), K, device=dev),
                  torch.randint(0, K, (B, N), device=dev, generator=g))
t = torch.full((B,), a.t, dtype=torch.int64, device=dev)
xp = torch.empty_like(x_t)
sc = head.head_scratch(B, N, dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.launches + 1)]
ev[0].record()
for i in range(a.launches):
    head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=2.0, seed=1, offset=i, x_prev_out=xp, scratch=sc, stats_1xtf32=not a.stats3x)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.launches)]
print("ms per launch:", ["%.3f" % x for x in ms], "token-updates/s: %.3e" % (B * N / min(ms) * 1e3), "redo rows", int(sc[1]))

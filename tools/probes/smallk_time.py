"""Stream kernel at the config-2 grid with smaller codebooks (K = 2048 is what the reference's UCF job actually uses)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3pm_b200 import _lib, ops
from oracle import d3pm_oracle as O
dev = "cuda:0"
T, N, B = 100, 4096, 16
for K, guidance in ((2048, True), (1024, True), (2048, False), (4096, True)):
    table = ops.build_coef_table(O.pack_schedule(O.make_schedule(T, K)).to(dev), T, K)
    g = torch.Generator(device=dev).manual_seed(0)
    lc = torch.randn(B, N, K, device=dev, generator=g)
    lu = torch.randn(B, N, K, device=dev, generator=g) if guidance else None
    x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < 0.5, torch.full((B, N), K, device=dev), torch.randint(0, K, (B, N), device=dev, generator=g))
    t = torch.full((B,), 50, dtype=torch.int64, device=dev)
    xp = torch.empty_like(x_t)
    def step(i):
        ops.fused_step(lc, lu, x_t, t, table, guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX, seed=1, offset=i, x_prev_out=xp)
    for i in range(5): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100): step(5 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    nbytes = B * N * ((2 if guidance else 1) * K * 4 + 16)
    print(f"K={K} guidance={guidance}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s")
    del lc, lu

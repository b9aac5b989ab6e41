"""Does the per-step time of the stream kernel drift under sustained load?  Chunks of 100 back-to-back steps."""
import os, sys, subprocess, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3pm_b200 import _lib, ops
from oracle import d3pm_oracle as O
dev = "cuda:0"
T, K, N, B = 100, 4096, 4096, 16
TNOW = int(os.environ.get("T_NOW", "50"))
sched = O.make_schedule(T, K)
table = ops.build_coef_table(O.pack_schedule(sched).to(dev), T, K)
g = torch.Generator(device=dev).manual_seed(0)
lc = torch.randn(B, N, K, device=dev, generator=g); lu = torch.randn(B, N, K, device=dev, generator=g)
pm = float(sched['log_cumprod_ct'][TNOW].exp())
x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < pm, torch.full((B, N), K, device=dev), torch.randint(0, K, (B, N), device=dev, generator=g))
t = torch.full((B,), TNOW, dtype=torch.int64, device=dev)
xp = torch.empty_like(x_t)
def q():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print("idle:", q())
for chunk in range(int(os.environ.get("CHUNKS", "20"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        ops.fused_step(lc, lu, x_t, t, table, guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX, seed=1, offset=chunk * 100 + i, x_prev_out=xp)
    e1.record(); torch.cuda.synchronize()
    print(f"chunk {chunk}: {e0.elapsed_time(e1) / 100:.4f} ms/step", q() if chunk % 5 == 4 else "")

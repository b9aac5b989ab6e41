"""Debug: per-group start / main-end / end timestamps of the stream kernel (library built with -DD3PM_STREAM_TIMING,
which turns the status word into a [groups][8] trace buffer).  D3PM_B200_LIB must point at that build."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from d3pm_b200 import _lib, ops
from oracle import d3pm_oracle as O
dev = "cuda:0"
T, K, N, B = 100, 4096, 4096, 16
table = ops.build_coef_table(O.pack_schedule(O.make_schedule(T, K)).to(dev), T, K)
g = torch.Generator(device=dev).manual_seed(0)
lc = torch.randn(B, N, K, device=dev, generator=g); lu = torch.randn(B, N, K, device=dev, generator=g)
PM = float(O.make_schedule(T, K)['log_cumprod_ct'][int(os.environ.get('T_NOW', '50'))].exp())
x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < PM, torch.full((B, N), K, device=dev), torch.randint(0, K, (B, N), device=dev, generator=g))
TNOW = int(os.environ.get("T_NOW", "50"))
t = torch.full((B,), TNOW, dtype=torch.int64, device=dev)
xp = torch.empty_like(x_t)
for it in range(6):
    st = torch.zeros(592 * 8, dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.fused_step(lc, lu, x_t, t, table, guidance_scale=2.0, sample_mode=_lib.SAMPLE_PHILOX, seed=1, offset=it, x_prev_out=xp, status=st)
    e1.record(); torch.cuda.synchronize()
    a = st.cpu().numpy().view(np.uint32).reshape(592, 8).astype(np.int64)
    start = a[:, 0] + (a[:, 1] << 32); start -= start.min()
    main, end = a[:, 2], a[:, 3]
    fin = start + end
    print(f"launch {it}: event {e0.elapsed_time(e1)*1e3:.1f} us | start skew max {start.max()/1e3:.1f} us | main dur min/med/max {main.min()/1e3:.1f}/{np.median(main)/1e3:.1f}/{main.max()/1e3:.1f} | "
          f"end-of-main (abs) min/max {(start+main).min()/1e3:.1f}/{(start+main).max()/1e3:.1f} | finish min/med/max {fin.min()/1e3:.1f}/{np.median(fin)/1e3:.1f}/{fin.max()/1e3:.1f} | redo rows max {a[:,5].max()} total {a[:,5].sum()}")
    if it == 5:
        sm = a[:, 6]
        per_sm = {}
        for s_, m_ in zip(sm, main): per_sm.setdefault(int(s_), []).append(m_)
        v = sorted((np.mean(x), s_) for s_, x in per_sm.items())
        print("fastest SMs (main us, smid):", [(round(x/1e3,1), s_) for x, s_ in v[:8]])
        print("slowest SMs:", [(round(x/1e3,1), s_) for x, s_ in v[-8:]])
        # redo duration for groups with redo
        tail = end - main
        for r in range(0, 4):
            sel = a[:, 5] == r
            if sel.any(): print(f"groups with {r} redo rows: {sel.sum()}, tail us mean {tail[sel].mean()/1e3:.1f} max {tail[sel].max()/1e3:.1f}")

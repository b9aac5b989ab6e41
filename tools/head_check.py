"""Development check of the fused head kernel (d3pm_head_step) on a GPU: logits vs torch fp32, tokens vs the unfused path."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import d3pm_b200  # noqa: E402
from d3pm_b200 import _lib, head, ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
B, N, K, D, T = int(os.environ.get("B", 2)), int(os.environ.get("N", 300)), 4096, 64, 100
g = torch.Generator(device=dev).manual_seed(0)
ln = torch.nn.LayerNorm(D).to(dev)
lin = torch.nn.Linear(D, K).to(dev)
with torch.no_grad():
    ln.weight.copy_(1 + 0.1 * torch.randn(D, device=dev, generator=g))
    ln.bias.copy_(0.1 * torch.randn(D, device=dev, generator=g))
    lin.weight.mul_(4.0)
hc = torch.randn(B, N, D, device=dev, generator=g) * 2 + 0.3
hu = torch.randn(B, N, D, device=dev, generator=g)
hw = head.HeadWeights.from_module(torch.nn.Sequential(ln, lin))
print("logit bound", hw.logit_bound, "valid", hw.valid)
s = 2.0
with torch.no_grad():
    lc, lu = lin(ln(hc)), lin(ln(hu))
    want = s * lc.double() + (1 - s) * lu.double()
got = head.head_step(hw, hc, hu, None, None, None, guidance_scale=s, mode=_lib.HEAD_LOGITS)
torch.cuda.synchronize()
err = (got.double() - want).abs()
print("LOGITS max|err|", float(err.max()), "mean", float(err.mean()), "ref absmax", float(want.abs().max()))
bad = (err > 1e-3).nonzero()
if len(bad):
    print("first bad", bad[:10].tolist(), got[tuple(bad[0])].item(), want[tuple(bad[0])].item())

class _Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.content_emb = type("E", (), {"num_embed": K + 1})()
model = d3pm_b200.FusedDiffusionTransformer(transformer=_Stub(), diffusion_step=T, alpha_init_type="alpha1", guidance_scale=s,
                                            content_seq_len=N).to(dev)
table = model.coef_table()
for tv in (50, 3, 0, 99):
    p_mask = float(model.log_cumprod_ct[tv].exp())
    x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < p_mask, torch.full((B, N), K, device=dev),
                      torch.randint(0, K, (B, N), device=dev, generator=g))
    t = torch.full((B,), tv, dtype=torch.int64, device=dev)
    unf = ops.fused_step(lc.contiguous(), lu.contiguous(), x_t, t, table, guidance_scale=s, sample_mode=_lib.SAMPLE_PHILOX_EXACT,
                         seed=7, offset=3, want_gap=True)
    ref = head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=s, mode=_lib.HEAD_REFERENCE, seed=7, offset=3)
    fus = head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=s, mode=_lib.HEAD_STEP, seed=7, offset=3)
    torch.cuda.synchronize()
    near = unf["gap"] < 2e-4
    d1 = (ref != unf["x_prev"])
    d2 = (fus != unf["x_prev"])
    print(f"t={tv}: reference-vs-unfused diff {int(d1.sum())} (away from near-ties {int((d1 & ~near).sum())}); "
          f"fused-vs-unfused diff {int(d2.sum())} (away {int((d2 & ~near).sum())}); near-ties {int(near.sum())}; "
          f"fused-vs-reference {int((fus != ref).sum())}")
if os.environ.get("BENCH"):
    B, N = 16, 4096
    hc = torch.randn(B, N, D, device=dev, generator=g)
    hu = torch.randn(B, N, D, device=dev, generator=g)
    x_t = torch.full((B, N), K, device=dev)
    t = torch.full((B,), 50, dtype=torch.int64, device=dev)
    xp = torch.empty_like(x_t)
    sc = head.head_scratch(B, N, dev)
    for mode, name in ((_lib.HEAD_STEP, "fused head step"),):
        for i in range(3):
            head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=s, mode=mode, seed=1, offset=i, x_prev_out=xp, scratch=sc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            head.head_step(hw, hc, hu, x_t, t, table, guidance_scale=s, mode=mode, seed=1, offset=i, x_prev_out=xp, scratch=sc)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 20:.3f} ms per step (16 x 4096 tokens x 4096 classes), redo rows {int(sc[1])}")
    with torch.no_grad():
        for i in range(2):
            a, b = lin(ln(hc)), lin(ln(hu))
        e0.record()
        for i in range(5):
            a, b = lin(ln(hc)), lin(ln(hu))
            ops.fused_step(a, b, x_t, t, table, guidance_scale=s, sample_mode=_lib.SAMPLE_PHILOX, seed=1, offset=i, x_prev_out=xp)
        e1.record()
        torch.cuda.synchronize()
        print(f"unfused (torch LayerNorm+Linear fp32 x2, then d3pm_fused_step): {e0.elapsed_time(e1) / 5:.3f} ms per step")

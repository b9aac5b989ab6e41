"""BASELINE config 3: the whole T-step sampling chain through `FusedDiffusionTransformer.sample()`.

    python tools/loop_bench.py [--videos 16] [--steps 100] [--grid 16 16 16]

The reference's `Text2ImageTransformer` does not travel to the GPU box (and its attention materialises a
17 GB [B,nh,N,N] tensor per layer at N = 4096), so the denoiser here is a stand-in with the reference's
interface and output layout: token + position embedding (n_embd 64) -> LayerNorm -> Linear(64 -> 4096), i.e.
the reference's `to_logits` head (transformer_utils.py:352-356) on top of an embedding, returning a
`[B, K, N]` permuted view of `[B, N, K]` (transformer_utils.py:442-443).  What is measured is the loop
itself: two denoiser forwards + one fused update per timestep, tokens carried as int64, no host syncs
inside the loop, one status check at the end.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import d3pm_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=16)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--grid", type=int, nargs=3, default=[16, 16, 16])
ap.add_argument("--codes", type=int, default=4096)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, K, T = a.videos, a.codes, a.steps
N = a.grid[0] * a.grid[1] * a.grid[2]


class HeadDenoiser(torch.nn.Module):
    def __init__(self, K, N, n_embd=64):
        super().__init__()
        self.content_emb = torch.nn.Embedding(K + 1, n_embd)
        self.content_emb.num_embed = K + 1
        self.pos = torch.nn.Parameter(torch.randn(N, n_embd) * 0.02)
        self.time = torch.nn.Embedding(T, n_embd)
        self.to_logits = torch.nn.Sequential(torch.nn.LayerNorm(n_embd), torch.nn.Linear(n_embd, K))

    def hidden_states(self, x_t, cond, t):
        return self.content_emb(x_t) + self.pos + self.time(t)[:, None, :] + cond.mean(-1, keepdim=True)

    def forward(self, x_t, cond, t):
        return self.to_logits(self.hidden_states(x_t, cond, t)).permute(0, 2, 1)  # [B, K, N] view of [B, N, K]


torch.manual_seed(0)
den = HeadDenoiser(K, N).to(dev)
model = d3pm_b200.FusedDiffusionTransformer(transformer=den, diffusion_step=T, alpha_init_type="alpha1",
                                            guidance_scale=2.0, content_seq_len=N).to(dev)
cond, cf = torch.randn(B, 1, 512, device=dev), torch.zeros(B, 1, 512, device=dev)

def chain():
    model.manual_seed(1).sample(["x"] * B, None, cond, cf, filter_ratio=0)  # warm-up (allocator, table)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.manual_seed(2).sample(["x"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    assert out.shape == (B, N) and int(out.max()) < K, "a finished chain holds no [MASK]"
    return total, out


total, out = chain()
if K in (1024, 2048, 4096):  # the head folded into the update kernel (logits never written)
    model.enable_fused_head()
    assert model.fused_head_active
    total_f, out_f = chain()
    model.enable_fused_head(False)
    print(f"chain with the fused head (d3pm_head_step): {total_f * 1e3:.1f} ms total, {total_f / T * 1e3:.3f} ms/step, "
          f"{B * N * T / total_f / 1e6:.1f} M token-updates/s through sample(); first-step agreement with the unfused chain "
          f"is covered by tests/test_gpu_head.py (chains diverge after the first near-tie)")

# the update alone on the same shapes, for the split
x = torch.full((B, N), K, dtype=torch.int64, device=dev)
tt = torch.full((B,), T // 2, dtype=torch.int64, device=dev)
lc = den(x, cond, tt).permute(0, 2, 1)
lu = den(x, cf, tt).permute(0, 2, 1)
xp = torch.empty_like(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    d3pm_b200.ops.fused_step(lc, lu, x, tt, model.coef_table(), guidance_scale=2.0, sample_mode=2, seed=1, x_prev_out=xp)
e0.record()
for i in range(20):
    d3pm_b200.ops.fused_step(lc, lu, x, tt, model.coef_table(), guidance_scale=2.0, sample_mode=2, seed=1, offset=i, x_prev_out=xp)
e1.record()
torch.cuda.synchronize()
upd = e0.elapsed_time(e1) / 20
print(f"chain: {T} steps x {B} videos x {N} tokens x {K}+1 classes: {total * 1e3:.1f} ms total, "
      f"{total / T * 1e3:.3f} ms/step; fused update alone {upd:.3f} ms/step = {100 * upd * T / (total * 1e3):.1f}% of the chain; "
      f"{B * N * T / total / 1e6:.1f} M token-updates/s through sample()")

"""Time of one purity-prior step (prior_rule 2) at the config-2 shape, by sampling mode of the candidate draw."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import d3pm_b200
from d3pm_b200 import _lib, ops
dev = torch.device("cuda", 0)
T, K, N, B = 100, 4096, 4096, 16
class _Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.content_emb = type("E", (), {"num_embed": K + 1})()
m = d3pm_b200.FusedDiffusionTransformer(transformer=_Stub(), diffusion_step=T, alpha_init_type="alpha1", guidance_scale=2.0, content_seq_len=N).to(dev)
table = m.coef_table()
g = torch.Generator(device=dev).manual_seed(0)
lc = torch.randn(B, N, K, device=dev, generator=g); lu = torch.randn(B, N, K, device=dev, generator=g)
x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < 0.5, torch.full((B, N), K, device=dev), torch.randint(0, K, (B, N), device=dev, generator=g))
t = torch.full((B,), 50, dtype=torch.int64, device=dev)
def timed(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
outs = {}
for name, mode in (("PHILOX_EXACT", _lib.SAMPLE_PHILOX_EXACT), ("PHILOX", _lib.SAMPLE_PHILOX)):
    def f():
        outs[name] = ops.fused_step(lc, lu, x_t, t, table, guidance_scale=2.0, sample_mode=mode, seed=3, offset=7, sample_from=_lib.FROM_RECON, want_score=True)
    print(name, "candidate draw + purity: %.3f ms" % timed(f))
print("same candidates:", torch.equal(outs["PHILOX"]["x_prev"], outs["PHILOX_EXACT"]["x_prev"]), "same score:", torch.equal(outs["PHILOX"]["score"], outs["PHILOX_EXACT"]["score"]))
cand, score = outs["PHILOX"]["x_prev"], outs["PHILOX"]["score"]
nrev = torch.full((B,), 40, dtype=torch.int32, device=dev)
print("purity_select: %.3f ms" % timed(lambda: ops.purity_select(x_t, cand, score, nrev, K, seed=1, offset=2)))

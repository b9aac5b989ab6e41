"""Every public method of the drop-in class once, at the config-2 shape (16 videos x 4096 tokens x 4096 codes, 1.07 GB per [B,K+1,N] tensor; VIDEOS=4 for a smaller run):
wall time per call (CUDA events) - and, under ncu, the kernels each one launches."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import d3pm_b200
from d3pm_b200 import ops
dev = torch.device("cuda", 0)
T, K, N, B = 100, 4096, 4096, int(os.environ.get('VIDEOS', '16'))
g = torch.Generator(device=dev).manual_seed(0)
LC = torch.randn(B, N, K, device=dev, generator=g); LU = torch.randn(B, N, K, device=dev, generator=g)
class _Emb:
    num_embed = K + 1
class _Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.content_emb = _Emb()
        self.to_logits = torch.nn.Sequential(torch.nn.LayerNorm(64), torch.nn.Linear(64, K))
    def forward(self, x_t, cond, t):
        return (LC if cond is COND else LU).permute(0, 2, 1)  # no host sync: identity of the conditioning tensor
m = d3pm_b200.FusedDiffusionTransformer(transformer=_Stub(), diffusion_step=T, alpha_init_type="alpha1", guidance_scale=2.0, content_seq_len=N).to(dev)
COND = torch.ones(B, 1, 512, device=dev)
cond, cf = COND, -torch.ones(B, 1, 512, device=dev)
t = torch.full((B,), 50, dtype=torch.int64, device=dev)
x_t = torch.where(torch.rand(B, N, device=dev, generator=g) < 0.5, torch.full((B, N), K, device=dev), torch.randint(0, K, (B, N), device=dev, generator=g))
log_x = m.index_to_log_onehot(x_t, K + 1) if hasattr(m, "index_to_log_onehot") else ops.as_logical(ops.tokens_to_log_onehot_rows(x_t, K + 1), K + 1)
recon = m.cf_predict_start(log_x, cond, cf, t)
calls = {
    "p_sample": lambda: m.p_sample(log_x, cond, cf, t, [0] * B, 10),
    "p_pred": lambda: m.p_pred(log_x, cond, cf, t),
    "cf_predict_start": lambda: m.cf_predict_start(log_x, cond, cf, t),
    "predict_start": lambda: m.predict_start(log_x, cond, t),
    "q_posterior": lambda: m.q_posterior(recon, log_x, t),
    "log_sample_categorical": lambda: m.log_sample_categorical(recon),
    "q_pred": lambda: m.q_pred(log_x, t),
    "q_sample": lambda: m.q_sample(log_x, t),
    "p_sample_tokens (what sample() calls)": lambda: m.p_sample_tokens(x_t, cond, cf, t),
}
reps = int(os.environ.get("REPS", "5"))
for name, fn in calls.items():
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push(name)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print(f"{name:40s} {e0.elapsed_time(e1) / reps:8.3f} ms")

"""Time of one training step through the module API (forward + backward; the denoiser is a stub returning fixed logits that
require grad), at the shipped training shape 16 x 1024 tokens x 4096 codes.  Under ncu it yields the launch list."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import d3pm_b200
dev = torch.device("cuda", 0)
T, K, N, B = 100, 4096, 1024, 16
class _Emb:
    num_embed = K + 1
class _Stub(torch.nn.Module):
    def __init__(self, logits):
        super().__init__()
        self.content_emb = _Emb()
        self.to_logits = torch.nn.Sequential(torch.nn.LayerNorm(64), torch.nn.Linear(64, K))
        self.logits = logits
    def forward(self, x_t, cond, t):
        return self.logits.permute(0, 2, 1)
g = torch.Generator(device=dev).manual_seed(0)
logits = torch.randn(B, N, K, device=dev, generator=g, requires_grad=True)
m = d3pm_b200.FusedDiffusionTransformer(transformer=_Stub(logits), diffusion_step=T, alpha_init_type="alpha1", guidance_scale=2.0,
                                        content_seq_len=N, auxiliary_loss_weight=5e-4, adaptive_auxiliary_loss=True,
                                        mask_weight=[1, 1]).to(dev)
x0 = torch.randint(0, K, (B, N), device=dev, generator=g)
batch = {"content_token": x0, "condition_embed_token": torch.ones(B, 1, 512, device=dev)}
def step():
    logits.grad = None
    out = m(batch, return_loss=True, return_logits=False)
    out["loss"].backward()
for _ in range(3): step()
torch.cuda.synchronize()
reps = int(os.environ.get("REPS", "20"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): step()
e1.record(); torch.cuda.synchronize()
print("module training step (forward + backward, stub denoiser): %.3f ms" % (e0.elapsed_time(e1) / reps))

"""Training-side use of the path on the GPU: forward-process operators, the fused
variational-bound loss and its gradient w.r.t. the denoiser logits, against the reference's golden vectors and
the oracle (whose gradient is PyTorch autograd through the reference's op sequence).  Needs a B200."""
import glob
import types

import numpy as np
import pytest
import torch

import d3pm_b200
from d3pm_b200 import ops
from oracle import d3pm_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _lg(a):  # token-major numpy -> logical [B,C,N] tensor
    return torch.from_numpy(a).permute(0, 2, 1)


class LeafDenoiser(torch.nn.Module):
    """Returns a leaf logits tensor the way Text2ImageTransformer returns its output (a [B,K,N] view of [B,N,K])."""

    def __init__(self, K, logits_bnk):
        super().__init__()
        self.content_emb = types.SimpleNamespace(num_embed=K + 1)
        self.to_logits = torch.nn.Sequential(torch.nn.Identity(), torch.nn.Linear(1, 1))
        self.logits = logits_bnk

    def forward(self, x_t, cond, t):
        return self.logits.permute(0, 2, 1)


def _model(K, T, N, logits, aux=0.0, adaptive=False):
    return d3pm_b200.FusedDiffusionTransformer(transformer=LeafDenoiser(K, logits), diffusion_step=T, alpha_init_type="alpha1",
                                               guidance_scale=2.0, content_seq_len=N, auxiliary_loss_weight=aux,
                                               adaptive_auxiliary_loss=adaptive).to(DEV)


def test_forward_process_operators_against_golden():
    fx = H.load(f"{H.GOLDEN}/forward_process_k64.npz")
    T, K = int(fx["T"]), int(fx["K"])
    B, N = fx["x0"].shape
    m = _model(K, T, N, torch.zeros(B, N, K, device=DEV))
    t = torch.from_numpy(fx["t"]).to(DEV)
    soft = _lg(fx["soft"]).contiguous().to(DEV)                                  # reference-contiguous layout
    hot = O.index_to_log_onehot(torch.from_numpy(fx["x0"]), K + 1).to(DEV)       # the reference's one-hot layout
    hot_t = O.index_to_log_onehot(torch.from_numpy(fx["x_t"]), K + 1).to(DEV)
    for got, want in ((m.q_pred(soft, t), "qpred_soft"), (m.q_pred(hot, t), "qpred_hot"), (m.q_pred(hot, t - 1), "qpred_hot_tm1"),
                      (m.q_pred_one_timestep(hot_t, t), "qone_hot"), (m.q_pred_one_timestep(soft, t), "qone_soft")):
        w = _lg(fx[want])
        finite = torch.isfinite(w)
        assert torch.equal(torch.isfinite(got.cpu()), finite), want
        assert (got.cpu()[finite] - w[finite]).abs().max() <= 1e-5, want
    u = _lg(fx["uniform"]).contiguous()
    m.inject_uniform = lambda shape, dev: u.to(dev)
    xs = m.q_sample(hot, t).argmax(1).cpu().numpy()
    assert np.array_equal(xs, fx["q_sample"])
    assert np.array_equal(m.q_sample_tokens(torch.from_numpy(fx["x0"]).to(DEV), t).cpu().numpy(), fx["q_sample"])


@pytest.mark.parametrize("path", sorted(glob.glob(f"{H.GOLDEN}/train_loss_*.npz")), ids=lambda p: p.split("train_loss_")[-1][:-4])
def test_train_loss_and_gradient_against_golden(path):
    fx = H.load(path)
    T, K = int(fx["T"]), int(fx["K"])
    B, N = fx["x0"].shape
    logits = torch.from_numpy(fx["logits"]).to(DEV).requires_grad_(True)
    m = _model(K, T, N, logits, aux=float(fx["aux"]), adaptive=bool(fx["adaptive"]))
    t, pt = torch.from_numpy(fx["t"]).to(DEV), torch.from_numpy(fx["pt"]).to(DEV)
    m.sample_time = lambda b, device, method="uniform": (t, pt)
    u = _lg(fx["uniform"]).contiguous()
    m.inject_uniform = lambda shape, dev: u.to(dev)
    x0 = torch.from_numpy(fx["x0"]).to(DEV)
    out = m({"content_token": x0, "condition_embed_token": torch.ones(B, 1, 512, device=DEV)}, return_loss=True)
    assert abs(float(out["loss"]) - float(fx["loss"])) <= 2e-5 * abs(float(fx["loss"]))
    assert np.array_equal(out["pred_data"].cpu().numpy(), fx["x0_recon"])
    assert (out["logits"].cpu() - _lg(fx["probs"])).abs().max() <= 1e-4
    out["loss"].backward()
    g, want = logits.grad.cpu().numpy(), fx["grad_logits"]
    # (the float64 closed form itself differs from the fp32 reference by 2.3e-5 relative on the stress fixture)
    assert np.abs(g - want).max() <= 5e-5 * np.abs(want).max() + 1e-8
    # the per-video losses and the posterior, through _train_loss itself
    lmp, vb, x0r = m._train_loss(x0, torch.ones(B, 1, 512, device=DEV))
    assert np.abs(vb.detach().cpu().numpy() - fx["vb_loss"]).max() <= 2e-5 * np.abs(fx["vb_loss"]).max()
    assert (lmp.cpu() - _lg(fx["log_model_prob"])).abs().max() <= H.POST_TOL
    # bookkeeping of the importance sampler (:434-438): two calls so far with these t
    assert torch.equal(m.Lt_count.cpu()[fx["t"]], torch.full((B,), 2.0)) and float(m.Lt_count.sum()) == 2 * B
    m.check_status()


@pytest.mark.parametrize("B,N,K,tvals,scale,aux,seed", [
    (2, 16, 4096, [37, 0], 1.0, 5e-4, 500),
    (3, 8, 4096, [99, 1, 50], 8.0, 0.0, 510),
    (2, 12, 1000, [5, 70], 2.0, 1e-3, 520),
])
def test_train_loss_and_gradient_against_oracle(B, N, K, tvals, scale, aux, seed):
    T = 100
    sched = O.make_schedule(T, K)
    lc, _, _, _, u = O.synth_inputs(B, N, K, 50, sched, seed=seed, scale=scale, spikes=scale > 1)
    x0 = torch.randint(0, K, (B, N), generator=torch.Generator().manual_seed(seed + 1))
    t = torch.tensor(tvals)
    pt = torch.rand(B, generator=torch.Generator().manual_seed(seed + 2)) * 0.02 + 0.001
    ref_logits = lc.clone().requires_grad_(True)
    _, vb_o, x0r_o, xt_o, _ = O.train_loss(sched, ref_logits.permute(0, 2, 1), x0, t, pt, u, auxiliary_loss_weight=aux,
                                          adaptive_auxiliary_loss=True)
    (vb_o.sum() / (B * N)).backward()

    logits = lc.clone().to(DEV).requires_grad_(True)
    m = _model(K, T, N, logits, aux=aux, adaptive=True)
    m.sample_time = lambda b, device, method="uniform": (t.to(DEV), pt.to(DEV))
    m.inject_uniform = lambda shape, dev: u.to(dev)
    out = m({"content_token": x0.to(DEV), "condition_embed_token": torch.ones(B, 1, 512, device=DEV)}, return_loss=True,
            return_logits=False)
    assert "logits" not in out and np.array_equal(out["pred_data"].cpu().numpy(), x0r_o.numpy())
    want_loss = float(vb_o.sum() / (B * N))
    assert abs(float(out["loss"]) - want_loss) <= 2e-5 * abs(want_loss)
    out["loss"].backward()
    gw = ref_logits.grad.numpy()
    assert np.abs(logits.grad.cpu().numpy() - gw).max() <= 5e-5 * np.abs(gw).max() + 1e-8


def test_unsupported_training_inputs_fail_loudly():
    K, T, B, N = 64, 100, 2, 8
    logits = torch.zeros(B, N, K, device=DEV, requires_grad=True)
    m = _model(K, T, N, logits)
    with pytest.raises(NotImplementedError):
        m({"content_token": torch.zeros(B, N, dtype=torch.long, device=DEV)}, return_loss=True, is_train=False)
    m.sample_time = lambda b, device, method="uniform": (torch.full((B,), 5, device=DEV), torch.full((B,), 0.01, device=DEV))
    bad = torch.full((B, N), K + 3, dtype=torch.long, device=DEV)  # clean tokens must be codes
    with pytest.raises(AssertionError):  # _train_loss reads the status word next to its own host syncs (the reference asserts, :45-46)
        m({"content_token": bad}, return_loss=True, return_logits=False)
    m.check_status()  # the word was cleared by the failed call: a stale bit must not fail a later, unrelated call


@pytest.mark.parametrize("K,scale", [(4096, 1.0), (4096, 9.0), (1024, 3.0)])
def test_stream_kernel_equals_row_kernel(K, scale):
    """The persistent TMA-pipelined training kernel (>= 2048 rows, K in {1024, 2048, 4096}: losses and gradient in one
    pass) against the one-CTA-per-row kernel (which the fixtures above pin to the reference), video by video."""
    from d3pm_b200 import train
    T, B, N = 100, 3, 1024
    g = torch.Generator().manual_seed(K + int(scale))
    logits = (torch.randn(B, N, K, generator=g) * scale)
    if scale > 5:  # peaked rows: both clamps fire
        logits[:, ::7, 5] += 200.0
    logits = logits.to(DEV)
    x0 = torch.randint(0, K, (B, N), generator=g).to(DEV)
    t = torch.tensor([0, 37, 99]).to(DEV)
    m = _model(K, T, N, logits)
    x_t = m.q_sample_tokens(x0, t)
    x_t[1, :50] = x0[1, :50]  # unmasked positions that kept their token
    table = m.coef_table()
    w_main = torch.tensor([0.7, 1.3, 0.01], device=DEV)
    w_aux = torch.tensor([0.2, 0.0, 0.5], device=DEV)
    both = train._train_rows(logits, K, x0, x_t, t, table, (1.0, 0.5), backward=2, w_main=w_main, w_aux=w_aux, want_recon=True)
    for b in range(B):  # one video = 1024 rows: below the stream kernel's threshold, so these run the row kernel
        sl = slice(b, b + 1)
        f = train._train_rows(logits[sl].contiguous(), K, x0[sl].contiguous(), x_t[sl].contiguous(), t[sl].contiguous(), table,
                              (1.0, 0.5), backward=0, want_recon=True)
        gr = train._train_rows(logits[sl].contiguous(), K, x0[sl].contiguous(), x_t[sl].contiguous(), t[sl].contiguous(), table,
                               (1.0, 0.5), backward=1, w_main=w_main[sl].contiguous(), w_aux=w_aux[sl].contiguous())
        for name in ("tok_main", "tok_aux"):
            a, w = both[name][sl], f[name]
            assert (a - w).abs().max() <= 2e-5 * w.abs().max() + 1e-6, (name, b)
        assert torch.equal(both["x0_recon"][sl], f["x0_recon"])
        assert (both["xtm1_recon"][sl] != f["xtm1_recon"]).float().mean() <= 0.002  # exact ties of clamped classes aside
        gw = gr["grad"]
        assert (both["grad"][sl] - gw).abs().max() <= 3e-5 * gw.abs().max() + 1e-9, b
    fwd = train._train_rows(logits, K, x0, x_t, t, table, (1.0, 0.5), backward=0, want_recon=True)
    assert torch.equal(fwd["tok_main"], both["tok_main"]) and torch.equal(fwd["x0_recon"], both["x0_recon"])
    only = train._train_rows(logits, K, x0, x_t, t, table, (1.0, 0.5), backward=1, w_main=w_main, w_aux=w_aux)
    assert torch.equal(only["grad"], both["grad"])


def test_upstream_gradient_other_than_announced():
    """`forward` announces d loss / d vb = 1 / (B N); a caller that scales the loss afterwards still gets the right gradient
    (d3pm_scale_rows), and one that uses vb_loss directly goes through the same route."""
    T, K, B, N = 100, 1024, 2, 1024
    g = torch.Generator().manual_seed(3)
    base = torch.randn(B, N, K, generator=g).to(DEV)
    x0 = torch.randint(0, K, (B, N), generator=g).to(DEV)
    tt, pt = torch.tensor([20, 70], device=DEV), torch.tensor([0.01, 0.02], device=DEV)
    grads = []
    for factor in (1.0, 3.0):
        logits = base.clone().requires_grad_(True)
        m = _model(K, T, N, logits, aux=1e-3)
        m.sample_time = lambda b, device, method="uniform": (tt, pt)
        m.manual_seed(5)
        out = m({"content_token": x0, "condition_embed_token": torch.ones(B, 1, 512, device=DEV)}, return_loss=True, return_logits=False)
        (out["loss"] * factor).backward()
        grads.append(logits.grad.clone())
    assert (grads[1] - 3.0 * grads[0]).abs().max() <= 1e-6 * grads[0].abs().max()
    logits = base.clone().requires_grad_(True)
    m = _model(K, T, N, logits, aux=1e-3)
    m.sample_time = lambda b, device, method="uniform": (tt, pt)
    m.manual_seed(5)
    _, vb, _ = m._train_loss(x0, torch.ones(B, 1, 512, device=DEV), need_log_model_prob=False)
    (vb * torch.tensor([1.0, 0.25], device=DEV)).sum().backward()
    want = grads[0] * (B * N)
    want[1] *= 0.25
    assert (logits.grad - want).abs().max() <= 2e-6 * want.abs().max()


@pytest.mark.parametrize("K,N", [(4096, 1024), (2048, 700), (64, 333)])
def test_fused_q_sample_tokens_equals_the_three_kernel_route(K, N):
    """d3pm_q_sample_tokens (one kernel, thinned race, no [B, K+1, N] tensor) draws exactly the tokens of
    tokens_to_log_onehot -> q_pred -> Gumbel-max with the same Philox stream, for every timestep incl. the wrap at
    t = -1 and for x_0 = [MASK]; the [MASK] rate follows the schedule."""
    from d3pm_b200 import _lib, ops, train
    T, B = 100, 12
    m = _model(K, T, N, torch.zeros(1, 1, K, device=DEV))
    g = torch.Generator(device=DEV).manual_seed(K + N)
    x0 = torch.randint(0, K, (B, N), device=DEV, generator=g)
    x0[0, :7] = K                                            # a few [MASK] inputs
    t = torch.tensor([0, 1, 5, 20, 37, 50, 63, 80, 95, 98, 99, -1], device=DEV)
    sched8 = m._sched8()
    kw = dict(seed=77, offset=5, row_offset=4321)
    fused = train.q_sample_tokens(x0, t, sched8, K, **kw)
    hot = ops.tokens_to_log_onehot_rows(x0, K + 1)
    qrows = train.q_pred_rows(hot, hot.shape[2], t, sched8, K, cumulative=True)
    ref = ops.gumbel_argmax_rows(qrows, qrows.shape[2], K + 1, noise_kind=2, **kw)
    torch.cuda.synchronize()
    assert torch.equal(fused, ref)
    ct = m.log_cumprod_ct[t[:-1]].exp().cpu()
    rate = (fused[1:-1] == K).float().mean(1).cpu()
    assert (rate - ct[1:]).abs().max() < 6.0 * (0.25 / N) ** 0.5

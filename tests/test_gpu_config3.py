"""BASELINE config 3 on the GPU: the drop-in class against the UNMODIFIED reference with its REAL denoiser.

The reference modules come from `baseline/reference_loader.py` (`/root/reference` in the build container, the staged
git-ignored copy `baseline/_ref/` on the GPU box); the tests skip only when neither is present.  The reference's
`Text2ImageTransformer.forward` calls `t.cuda()` (transformer_utils.py:439), so this can only run on a CUDA box.

* the reference `DiffusionTransformer`'s `state_dict()` loads strictly into `FusedDiffusionTransformer` and back;
* teacher-forced chain: at EVERY one of the T = 100 reverse steps the reference's own `p_sample` (its `torch.rand_like`
  replaced by a shared uniform tensor) and ours are fed the same `x_t`; posterior log-probs agree within 1e-4, tokens are
  identical except at logged near-ties, and the reference's `x_{t-1}` is what both see at the next step;
* a short chain at the config-3 token count N = 4096 (16 x 16 x 16 grid);
* free-running `sample()` of both classes ends without a [MASK] token, fused head included.
"""
import pytest
import torch

from baseline import reference_loader as RL
from oracle import d3pm_oracle as O
from tests import helpers as H
from tools import config3

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not RL.reference_available(), reason="reference not staged (baseline/_ref absent)")]
DEV = torch.device("cuda", 0)


def _models(B, N, K, T, side, gain=30.0):
    # random-init logits are ~N(0, 0.1): a gain on the head's weight makes p(x0 | x_t) as peaked as a trained model's
    return config3.build_models(B, N, K, T, [side, side], DEV, guidance=2.0, seed=0, logit_gain=gain)


def _teacher_forced(ref, ours, B, N, K, steps, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    cond = torch.randn(B, 77, 512, device=DEV, generator=g)
    cf = torch.randn(B, 77, 512, device=DEV, generator=g) * 0.1
    log_z = torch.log(torch.cat((torch.zeros(B, K, N, device=DEV), torch.ones(B, 1, N, device=DEV)), dim=1))
    worst, differ, near_total = 0.0, 0, 0
    with torch.no_grad():
        for ti in steps:
            t = torch.full((B,), ti, device=DEV, dtype=torch.long)
            u = torch.rand(B, K + 1, N, device=DEV, generator=g)
            post_ref, recon_ref = ref.p_pred(log_z, cond, cf, t)
            with RL.injected_uniform(lambda x: u):
                nxt_ref, _ = ref.p_sample(log_z, cond, cf, t, [0] * B, ref.n_sample[ti])
            tok_ref = nxt_ref.argmax(1)

            post, recon = ours.p_pred(log_z, cond, cf, t)
            ours.inject_uniform = lambda shape, dev: u
            nxt, _ = ours.p_sample(log_z, cond, cf, t, [0] * B, ours.n_sample[ti])
            ours.inject_uniform = None
            tok = nxt.argmax(1)

            worst = max(worst, float((post - post_ref).abs().max()), float((recon - recon_ref).abs().max()))
            near = O.near_ties(post_ref.cpu(), u.cpu()).numpy()
            diff = (tok != tok_ref).cpu().numpy()
            differ += int(diff.sum())
            near_total += int(near.sum())
            assert not (diff & ~near).any(), f"t={ti}: {int((diff & ~near).sum())} tokens differ away from near-ties"
            log_z = nxt_ref  # teacher forcing: both arms continue from the reference's state
    return worst, differ, near_total, log_z


def test_reference_state_dict_round_trip():
    ref, ours = _models(1, 64, 1024, 100, 8)
    sd = ref.state_dict()
    assert set(sd.keys()) == set(ours.state_dict().keys())
    for k, v in ours.state_dict().items():
        assert torch.equal(v, sd[k]), k
    ref.load_state_dict(ours.state_dict(), strict=True)  # and back: a checkpoint written by the drop-in loads upstream


def test_teacher_forced_chain_all_steps_n1024():
    B, N, K, T = 2, 1024, 4096, 100
    ref, ours = _models(B, N, K, T, 32)
    worst, differ, near, log_z = _teacher_forced(ref, ours, B, N, K, range(T - 1, -1, -1), seed=5)
    print(f"[config 3, N=1024] 100 teacher-forced steps: max |post/recon - reference| = {worst:.2e}; "
          f"{differ} tokens differ, all among {near} logged near-ties")
    assert worst <= H.POST_TOL
    assert int(log_z.argmax(1).max()) < K  # the reference's chain itself ended without [MASK]


def test_teacher_forced_chain_n4096():
    B, N, K, T = 1, 4096, 4096, 100
    ref, ours = _models(B, N, K, T, 64)
    worst, differ, near, _ = _teacher_forced(ref, ours, B, N, K, [99, 98, 97], seed=6)
    print(f"[config 3, N=4096] 3 teacher-forced steps: max |post/recon - reference| = {worst:.2e}; "
          f"{differ} tokens differ, all among {near} logged near-ties")
    assert worst <= H.POST_TOL


def test_free_running_sample_both_classes():
    B, N, K, T = 2, 256, 1024, 25
    ref, ours = _models(B, N, K, T, 16, gain=10.0)
    g = torch.Generator(device=DEV).manual_seed(3)
    cond, cf = torch.randn(B, 77, 512, device=DEV, generator=g), torch.zeros(B, 77, 512, device=DEV)
    a = ref.sample(["x"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    b = ours.manual_seed(4).sample(["x"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    ours.enable_fused_head()
    assert ours.fused_head_active
    c = ours.manual_seed(4).sample(["x"] * B, None, cond, cf, filter_ratio=0)["content_token"]
    for tok in (a, b, c):
        assert tok.shape == (B, N) and tok.dtype == torch.int64 and int(tok.max()) < K and int(tok.min()) >= 0


def test_identical_conditioning_runs_the_denoiser_once_and_changes_nothing():
    """The reference's pipeline zeroes both text embeddings (networks/discrete_diffusion.py:25, :49): its two denoiser passes
    per step then compute the same logits.  The drop-in runs ONE pass in that case; tokens and posterior are those of the
    reference's two-pass `p_sample` / `p_pred` on the same noise, and with the sharing switched off."""
    B, N, K, T = 2, 256, 1024, 100
    ref, ours = _models(B, N, K, T, 16)
    calls = {"n": 0}
    ours.transformer.register_forward_hook(lambda *a: calls.__setitem__("n", calls["n"] + 1))
    cond, cf = torch.zeros(B, 1, 512, device=DEV), torch.zeros(B, 1, 512, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(8)
    x = torch.where(torch.rand(B, N, device=DEV, generator=g) < 0.5, torch.full((B, N), K, device=DEV),
                    torch.randint(0, K, (B, N), device=DEV, generator=g))
    log_x = ref_index_to_log_onehot(x, K + 1)
    t = torch.tensor([60, 7], device=DEV)
    u = torch.rand(B, K + 1, N, device=DEV, generator=g)
    with torch.no_grad():
        post_ref, _ = ref.p_pred(log_x, cond, cf, t)
        with RL.injected_uniform(lambda z: u):
            nxt_ref, _ = ref.p_sample(log_x, cond, cf, t, [0] * B, ref.n_sample[60])
    calls["n"] = 0
    post, _ = ours.p_pred(log_x, cond, cf, t)
    assert calls["n"] == 1                               # one denoiser pass for both guidance branches
    ours.inject_uniform = lambda shape, dev: u
    nxt, _ = ours.p_sample(log_x, cond, cf, t, [0] * B, ours.n_sample[60])
    ours.inject_uniform = None
    assert float((post - post_ref).abs().max()) <= H.POST_TOL
    near = O.near_ties(post_ref.cpu(), u.cpu()).numpy()
    diff = (nxt.argmax(1) != nxt_ref.argmax(1)).cpu().numpy()
    assert not (diff & ~near).any()
    # the same drop-in with the sharing off: two passes, the very same tokens (Philox noise, same key)
    a = ours.manual_seed(3).p_sample_tokens(x, cond, cf, t)
    ours.share_identical_conditioning = False
    calls["n"] = 0
    b = ours.manual_seed(3).p_sample_tokens(x, cond, cf, t)
    assert calls["n"] == 2 and torch.equal(a, b)
    # different embeddings, or a denoiser in training mode (dropout): never shared
    ours.share_identical_conditioning = True
    calls["n"] = 0
    ours.p_sample_tokens(x, cond, cf + 1.0, t)
    assert calls["n"] == 2


def ref_index_to_log_onehot(x, C):
    return RL.load_diffusion_module().index_to_log_onehot(x, C)

"""CPU oracle for the D3PM / VQ-Diffusion reverse-diffusion token update.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product path (`d3pm_b200`) never does and fails
loudly when its CUDA library is missing.

It restates, with plain PyTorch CPU ops in the reference's own logical layout
``[B, K+1, N]`` (class dimension = 1), the algorithm of

    /root/reference/src/models/motionencoder/diffusion_transformer.py

Every function cites the reference lines it follows.  Parity pinning: the
reference holds no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by `tests/golden/make_golden.py` (which imports the reference by path)
and committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks the
oracle against them bit-for-bit (tokens) / to 1e-6 (log-probs).

Two layers live here:

* the *op-faithful* functions (`q_posterior`, `cf_predict_start_from_logits`,
  `log_sample_categorical`, `p_sample_step`) mirror the reference's sequence of
  tensor ops, so timing them is a fair "reference PyTorch CPU path" baseline;
* `closed_form_rows` is an independent float64 per-token derivation of the same
  quantities (the algebra the CUDA kernel uses), used to cross-check the kernel
  design without a GPU.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

LOG_TINY = math.log(1e-30)  # the reference's "log zero" for one-hot entries (:50, :258)
CLAMP_LO = -70.0            # the reference's probability floor in log space (:236, :247, :283)

SCHEDULE_NAMES = (
    "log_at", "log_bt", "log_ct", "log_1_min_ct",                        # length T
    "log_cumprod_at", "log_cumprod_bt", "log_cumprod_ct", "log_1_min_cumprod_ct",  # length T+1
)


# --------------------------------------------------------------------------- schedule
def make_schedule(num_timesteps: int, num_codes: int) -> Dict[str, torch.Tensor]:
    """Mask-and-replace schedule buffers ("alpha1" init).

    Follows `alpha_schedule` (diffusion_transformer.py:56-69) and the buffer
    construction in `DiffusionTransformer.__init__` (:115-149): linear cumulative
    keep / mask probabilities, per-step ratios, logs in float64, stored as float32.
    The cumulative rows get one extra trailing slot (index T) that encodes "t = -1"
    (keep everything), which `q_pred` reaches through its modulo (:203).
    """
    T, K = int(num_timesteps), int(num_codes)
    keep_first, keep_last = 0.99999, 0.000009
    mask_first, mask_last = 0.000009, 0.99999
    ramp = np.arange(0, T) / (T - 1)
    keep_cum = np.concatenate(([1.0], ramp * (keep_last - keep_first) + keep_first))
    mask_cum = np.concatenate(([0.0], ramp * (mask_last - mask_first) + mask_first))
    keep_step = keep_cum[1:] / keep_cum[:-1]
    stay_unmasked = (1 - mask_cum)[1:] / (1 - mask_cum)[:-1]
    mask_step = 1 - stay_unmasked
    repl_step = (1 - keep_step - mask_step) / K
    keep_cum = np.concatenate((keep_cum[1:], [1.0]))
    mask_cum = np.concatenate((mask_cum[1:], [0.0]))
    repl_cum = (1 - keep_cum - mask_cum) / K

    def lg(a):
        with np.errstate(divide="ignore"):
            return torch.log(torch.tensor(np.asarray(a, dtype=np.float64)))

    def log_one_minus(la):  # log_1_min_a (:29-30)
        return torch.log(1 - la.exp() + 1e-40)

    log_ct, log_cum_ct = lg(mask_step), lg(mask_cum)
    rows = {
        "log_at": lg(keep_step), "log_bt": lg(repl_step), "log_ct": log_ct,
        "log_1_min_ct": log_one_minus(log_ct),
        "log_cumprod_at": lg(keep_cum), "log_cumprod_bt": lg(repl_cum),
        "log_cumprod_ct": log_cum_ct, "log_1_min_cumprod_ct": log_one_minus(log_cum_ct),
    }
    return {k: v.float() for k, v in rows.items()}


def pack_schedule(sched: Dict[str, torch.Tensor]) -> torch.Tensor:
    """[8, T+1] float32 matrix in SCHEDULE_NAMES order (per-step rows zero-padded)."""
    T1 = sched["log_cumprod_at"].numel()
    out = torch.zeros(8, T1, dtype=torch.float32)
    for i, name in enumerate(SCHEDULE_NAMES):
        row = sched[name].detach().float().cpu()
        out[i, : row.numel()] = row
    return out


# --------------------------------------------------------------------------- helpers
def log_add_exp(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """log(exp(a)+exp(b)), max-shifted (:32-34)."""
    top = torch.maximum(a, b)
    return top + torch.log(torch.exp(a - top) + torch.exp(b - top))


def _per_batch(row: torch.Tensor, t: torch.Tensor, ndim: int) -> torch.Tensor:
    """`extract` (:36-39): gather schedule entries per batch element -> [B,1,1,...]."""
    return row.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))


def index_to_log_onehot(x: torch.Tensor, num_classes: int) -> torch.Tensor:
    """int64 [B,N] -> log one-hot, logically [B,C,N] (:44-51); zeros become log(1e-30)."""
    assert int(x.max()) < num_classes
    hot = torch.nn.functional.one_hot(x, num_classes)
    order = (0, x.dim()) + tuple(range(1, x.dim()))
    return torch.log(hot.permute(order).float().clamp(min=1e-30))


def log_onehot_to_index(log_x: torch.Tensor) -> torch.Tensor:
    """(:53-54)"""
    return log_x.argmax(1)


# --------------------------------------------------------------------------- forward process
def q_pred_one_timestep(sched, log_x_t: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """q(x_t | x_{t-1}) in log space (:185-199)."""
    nd = log_x_t.dim()
    la, lb = _per_batch(sched["log_at"], t, nd), _per_batch(sched["log_bt"], t, nd)
    lc, l1c = _per_batch(sched["log_ct"], t, nd), _per_batch(sched["log_1_min_ct"], t, nd)
    codes = log_add_exp(log_x_t[:, :-1, :] + la, lb)
    mask_row = log_add_exp(log_x_t[:, -1:, :] + l1c, lc)
    return torch.cat([codes, mask_row], dim=1)


def q_pred(sched, log_x_start: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """q(x_t | x_0) in log space with the cumulative schedule; t wraps mod T+1 (:201-218)."""
    T1 = sched["log_cumprod_at"].numel()
    t = (t + T1) % T1
    nd = log_x_start.dim()
    lA, lB = _per_batch(sched["log_cumprod_at"], t, nd), _per_batch(sched["log_cumprod_bt"], t, nd)
    lC, l1C = _per_batch(sched["log_cumprod_ct"], t, nd), _per_batch(sched["log_1_min_cumprod_ct"], t, nd)
    codes = log_add_exp(log_x_start[:, :-1, :] + lA, lB)
    mask_row = log_add_exp(log_x_start[:, -1:, :] + l1C, lC)
    return torch.cat([codes, mask_row], dim=1)


# --------------------------------------------------------------------------- denoiser post-processing
def predict_start_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """p(x_0 | x_t) from raw denoiser logits, logically [B,K,N] (:231-236).

    float64 log-softmax over classes, back to float32, a constant -70 row appended for
    the [MASK] class, everything clamped to [-70, 0].  (The denoiser call itself,
    :221-226, is outside the accelerated path.)
    """
    B, K, N = logits.shape
    lp = torch.log_softmax(logits.double(), dim=1).float()
    floor_row = torch.zeros(B, 1, N, dtype=lp.dtype) - 70
    return torch.clamp(torch.cat((lp, floor_row), dim=1), CLAMP_LO, 0)


def cf_predict_start_from_logits(logits_c: torch.Tensor, logits_u: Optional[torch.Tensor],
                                 guidance_scale: float) -> torch.Tensor:
    """Classifier-free guidance combine (:240-249).

    `logits_u is None` is "guidance off".  In the reference the |s-1|<1e-3 branch
    (:242-243) raises AttributeError, so the guidance-off result is `predict_start`
    of the conditional logits alone.
    """
    cond = predict_start_from_logits(logits_c)
    if logits_u is None:
        return cond
    B, _, N = logits_c.shape
    cond = cond[:, :-1]
    unc = predict_start_from_logits(logits_u)[:, :-1]
    mix = unc + guidance_scale * (cond - unc)
    mix -= torch.logsumexp(mix, dim=1, keepdim=True)
    mix = mix.clamp(CLAMP_LO, 0)
    floor_row = torch.zeros(B, 1, N, dtype=mix.dtype) - 70
    return torch.cat((mix, floor_row), dim=1)


# --------------------------------------------------------------------------- posterior
def q_posterior(sched, log_x_start: torch.Tensor, log_x_t: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """p_theta(x_{t-1} | x_t) = sum_x0 q(x_{t-1} | x_t, x0) p(x0 | x_t), log space (:251-283)."""
    T = sched["log_at"].numel()
    assert int(t.min()) >= 0 and int(t.max()) < T
    B, C, N = log_x_start.shape
    x_t = log_onehot_to_index(log_x_t)
    masked = (x_t == C - 1).unsqueeze(1)
    log_one = torch.zeros(B, 1, 1, dtype=log_x_t.dtype)
    log_tiny = torch.log(log_one + 1.0e-30).expand(-1, -1, N)

    # q(x_t | x_0) per candidate x_0; masked positions see the cumulative mask rate instead
    lqt = q_pred(sched, log_x_t, t)[:, :-1, :]
    cum_ct = _per_batch(sched["log_cumprod_ct"], t, 3).expand(-1, C - 1, -1)
    lqt = (~masked) * lqt + masked * cum_ct

    # q(x_t | x_{t-1}); mask row forced to "log zero", masked positions see c_t / 1
    lq1 = q_pred_one_timestep(sched, log_x_t, t)
    lq1 = torch.cat((lq1[:, :-1, :], log_tiny), dim=1)
    ct = _per_batch(sched["log_ct"], t, 3).expand(-1, C - 1, -1)
    ct = torch.cat((ct, log_one), dim=1)
    lq1 = (~masked) * lq1 + masked * ct

    q = log_x_start[:, :-1, :] - lqt
    q = torch.cat((q, log_tiny), dim=1)
    norm = torch.logsumexp(q, dim=1, keepdim=True)
    q = q - norm
    out = q_pred(sched, q, t - 1) + lq1 + norm
    return torch.clamp(out, CLAMP_LO, 0)


# --------------------------------------------------------------------------- sampler
def gumbel_from_uniform(u: torch.Tensor) -> torch.Tensor:
    """(:356)"""
    return -torch.log(-torch.log(u + 1e-30) + 1e-30)


def log_sample_categorical(logits: torch.Tensor, uniform: torch.Tensor,
                           return_index: bool = False):
    """Gumbel-max draw over dim 1 (:354-359) with the uniform noise injected.

    The reference draws `torch.rand_like(logits)`; here the same-shaped tensor is an
    argument so the CUDA path and the reference can share it.
    """
    idx = (gumbel_from_uniform(uniform) + logits).argmax(dim=1)
    if return_index:
        return idx
    return index_to_log_onehot(idx, logits.shape[1])


def p_sample_step(sched, logits_c: torch.Tensor, logits_u: Optional[torch.Tensor],
                  log_x_t: torch.Tensor, t: torch.Tensor, guidance_scale: float,
                  uniform: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One reverse step = p_pred (:285-296) + the prior_rule==0 branch of p_sample (:347-350).

    `logits_c` / `logits_u` are the denoiser outputs, logically [B,K,N].
    Returns (log one-hot of x_{t-1} [B,K+1,N], posterior log-probs, log_x_recon).
    """
    recon = cf_predict_start_from_logits(logits_c, logits_u, guidance_scale)
    post = q_posterior(sched, recon, log_x_t, t)
    return log_sample_categorical(post, uniform), post, recon


def near_ties(post: torch.Tensor, uniform: torch.Tensor, gap: float = 2e-4) -> torch.Tensor:
    """bool [B,N]: positions whose top-2 Gumbel scores are closer than `gap` (to be logged)."""
    top2 = (gumbel_from_uniform(uniform) + post).topk(2, dim=1).values
    return (top2[:, 0] - top2[:, 1]) < gap


# --------------------------------------------------------------------------- synthetic inputs (SURVEY.md §8 d)
def synth_inputs(B: int, N: int, K: int, t_value, sched, *, seed: int = 0, scale: float = 1.0,
                 spikes: bool = False):
    """Seeded CPU-generator inputs shared by the oracle, the golden generator and the GPU tests.

    Physical layout of the logits is token-major [B,N,K] like the denoiser's output;
    reference-layout consumers take `.permute(0,2,1)` views.
    """
    g = torch.Generator().manual_seed(seed)
    lc = torch.randn(B, N, K, generator=g) * scale
    g = torch.Generator().manual_seed(seed + 1)
    lu = torch.randn(B, N, K, generator=g) * scale
    if spikes:  # peaked rows that make every -70 clamp fire
        g = torch.Generator().manual_seed(seed + 4)
        rows = torch.randint(0, N, (max(1, N // 4),), generator=g)
        cols = torch.randint(0, K, (max(1, N // 4),), generator=g)
        lc[:, rows, cols] += 150.0
        lu[:, rows.flip(0), cols] += 90.0
    if torch.is_tensor(t_value):
        t = t_value.clone().long()
    else:
        t = torch.full((B,), int(t_value), dtype=torch.long)
    g = torch.Generator().manual_seed(seed + 2)
    p_mask = sched["log_cumprod_ct"][t].exp().view(B, 1)
    is_mask = torch.rand(B, N, generator=g) < p_mask
    codes = torch.randint(0, K, (B, N), generator=g)
    x_t = torch.where(is_mask, torch.full_like(codes, K), codes)
    g = torch.Generator().manual_seed(seed + 3)
    u = torch.rand(B, K + 1, N, generator=g)
    return lc, lu, x_t, t, u


# --------------------------------------------------------------------------- independent float64 derivation
def closed_form_rows(sched, logits_c: np.ndarray, logits_u: Optional[np.ndarray], x_t: np.ndarray,
                     t: np.ndarray, guidance_scale: float):
    """Per-token closed form of recon + posterior in float64 (token-major rows).

    logits_*: [B,N,K]; x_t: [B,N]; t: [B].  Returns (recon [B,N,K+1], post [B,N,K+1]).
    Derivation: with p_k = exp(recon_k), W_k = exp(-log q(x_t|x0=k)),
    e^L = sum_k p_k W_k + 1e-30,
        post_k = log(p_k W_k Abar' + Bbar' e^L) + log q(x_t | x_{t-1}=k)
        post_K = log(1e-30 (1 - Cbar') + Cbar' e^L) + one_K
    where the primed quantities are the cumulative schedule at t-1 (identity at t=0).
    The clamps of :236, :247 and :283 are applied in the same order as the reference.
    """
    S = {k: v.double().numpy() for k, v in sched.items()}
    T = S["log_at"].shape[0]
    B, N, K = logits_c.shape

    def lsm(x):
        x = x.astype(np.float64)
        m = x.max(-1, keepdims=True)
        return x - (m + np.log(np.exp(x - m).sum(-1, keepdims=True)))

    def lae(a, b):
        m = np.maximum(a, b)
        with np.errstate(invalid="ignore", divide="ignore"):
            r = m + np.log(np.exp(a - m) + np.exp(b - m))
        return np.where(np.isneginf(m), -np.inf, r)

    lc = np.clip(lsm(logits_c).astype(np.float32).astype(np.float64), CLAMP_LO, 0)
    if logits_u is None:
        rec = lc
    else:
        lu = np.clip(lsm(logits_u).astype(np.float32).astype(np.float64), CLAMP_LO, 0)
        y = lu + guidance_scale * (lc - lu)
        rec = np.clip(lsm(y), CLAMP_LO, 0)
    recon = np.concatenate([rec, np.full((B, N, 1), CLAMP_LO)], -1)

    post = np.empty((B, N, K + 1))
    Z = LOG_TINY
    for b in range(B):
        tb = int(t[b]); tp = (tb - 1) % (T + 1)
        la, lb_, lct = S["log_at"][tb], S["log_bt"][tb], S["log_ct"][tb]
        lA, lB, lC = S["log_cumprod_at"][tb], S["log_cumprod_bt"][tb], S["log_cumprod_ct"][tb]
        lAp, lBp = S["log_cumprod_at"][tp], S["log_cumprod_bt"][tp]
        lCp, l1Cp = S["log_cumprod_ct"][tp], S["log_1_min_cumprod_ct"][tp]
        with np.errstate(divide="ignore"):
            Ap, Bp, Cp, omCp = np.exp(lAp), np.exp(lBp), np.exp(lCp), np.exp(l1Cp)
        for n in range(N):
            j = int(x_t[b, n]); p = np.exp(rec[b, n])
            if j == K:
                W = np.full(K, np.exp(-lC)); one = np.full(K, lct); oneK = 0.0
            else:
                W = np.full(K, np.exp(-lae(Z + lA, lB))); W[j] = np.exp(-lae(lA, lB))
                one = np.full(K, lae(Z + la, lb_)); one[j] = lae(la, lb_); oneK = Z
            eL = (p * W).sum() + 1e-30
            with np.errstate(divide="ignore"):
                post[b, n, :K] = np.log(p * W * Ap + Bp * eL) + one
                post[b, n, K] = np.log(1e-30 * omCp + Cp * eL) + oneK
    return recon, np.clip(post, CLAMP_LO, 0)


# --------------------------------------------------------------------------- training side
def q_sample(sched, log_x_start: torch.Tensor, t: torch.Tensor, uniform: torch.Tensor) -> torch.Tensor:
    """Forward noising draw x_t ~ q(x_t | x_0) (:361-366): q_pred then the Gumbel-max sampler, noise injected."""
    return log_sample_categorical(q_pred(sched, log_x_start, t), uniform)


def multinomial_kl(log_p: torch.Tensor, log_q: torch.Tensor) -> torch.Tensor:
    """(:181-183)"""
    return (log_p.exp() * (log_p - log_q)).sum(dim=1)


def train_loss(sched, logits: torch.Tensor, x0: torch.Tensor, t: torch.Tensor, pt: torch.Tensor,
               uniform: torch.Tensor, *, auxiliary_loss_weight: float = 0.0, adaptive_auxiliary_loss: bool = False,
               mask_weight=(1, 1), is_train: bool = True):
    """The variational-bound training loss of `_train_loss` (:391-457) for given timesteps `t` (with their
    sampling probabilities `pt`), injected forward-noising noise and denoiser logits `[B, K, N]` (a leaf
    that may require grad; in the reference they are `transformer(x_t, cond, t)`).

    Returns (log_model_prob [B,K+1,N], vb_loss [B], x0_recon [B,N], x_t [B,N], kl_loss [B]).
    The running-average bookkeeping (`Lt_history`, `Lt_count`, acc lists; :407-417, :434-438) is the caller's.
    """
    T = sched["log_at"].numel()
    B, K, N = logits.shape
    C = K + 1
    log_x_start = index_to_log_onehot(x0, C)
    log_xt = q_sample(sched, log_x_start, t, uniform)
    xt = log_onehot_to_index(log_xt)

    log_x0_recon = predict_start_from_logits(logits)                      # P_theta(x0 | xt)        (:404)
    log_model_prob = q_posterior(sched, log_x0_recon, log_xt, t)          # through q(xt-1 | xt, x0) (:405)
    x0_recon = log_onehot_to_index(log_x0_recon)

    log_true_prob = q_posterior(sched, log_x_start, log_xt, t)            # (:420)
    kl = multinomial_kl(log_true_prob, log_model_prob)
    mask_region = (xt == C - 1).float()
    w = mask_region * mask_weight[0] + (1.0 - mask_region) * mask_weight[1]
    kl = (kl * w).reshape(B, -1).sum(-1)

    decoder_nll = -(log_x_start.exp() * log_model_prob).sum(dim=1)        # log_categorical (:41-42, :427)
    decoder_nll = decoder_nll.reshape(B, -1).sum(-1)
    at_zero = (t == torch.zeros_like(t)).float()
    kl_loss = at_zero * decoder_nll + (1.0 - at_zero) * kl
    vb_loss = kl_loss / pt
    if auxiliary_loss_weight != 0 and is_train:
        kl_aux = multinomial_kl(log_x_start[:, :-1, :], log_x0_recon[:, :-1, :])
        kl_aux = (kl_aux * w).reshape(B, -1).sum(-1)
        kl_aux_loss = at_zero * decoder_nll + (1.0 - at_zero) * kl_aux
        extra = (1 - t / T) + 1.0 if adaptive_auxiliary_loss else 1.0
        vb_loss = vb_loss + extra * auxiliary_loss_weight * kl_aux_loss / pt
    return log_model_prob, vb_loss, x0_recon, xt, kl_loss


# --------------------------------------------------------------------------- purity-prior sampling
def multinomial_without_replacement(weights: torch.Tensor, n: int, expo: torch.Tensor) -> torch.Tensor:
    """`torch.multinomial(weights, n)` (no replacement) with its noise injected.

    ATen draws `q ~ Exp(1)` per category and returns `topk(weights / q, n)` (argmax for n = 1)
    (aten/src/ATen/native/Distributions.cpp, multinomial_out); `tests/golden/make_golden_purity.py` checks that
    equivalence against the real op before generating fixtures.  `expo` is that `q`."""
    return torch.topk(weights / expo, n).indices


def p_sample_purity_step(sched, logits_c: torch.Tensor, logits_u: Optional[torch.Tensor], log_x_t: torch.Tensor,
                         t: torch.Tensor, guidance_scale: float, uniform: torch.Tensor, expo: torch.Tensor,
                         sampled, to_sample: int, *, prior_rule: int, prior_weight: float = 0.0, prior_ps: int = 1024):
    """`p_sample` with `prior_rule` 1 / 2 (:304-352): Improved-VQ-Diffusion "high-quality inference" / purity prior.

    uniform: `[B,K+1,N]` (the `rand_like` of the Gumbel draw), expo: `[B,N]` (the Exp(1) of each video's
    `torch.multinomial`).  Returns (x_{t-1} tokens `[B,N]`, updated `sampled` list).
    """
    recon = cf_predict_start_from_logits(logits_c, logits_u, guidance_scale)
    B, C, N = recon.shape
    sampled = list(sampled)
    if int(t[0]) > 0 and prior_rule > 0:
        x_idx = log_onehot_to_index(log_x_t)
        if prior_rule == 1:
            score = torch.ones(B, N)
        else:
            score = torch.exp(recon).max(dim=1).values.clamp(0, 1)
            score = score / (score.max(dim=1, keepdim=True).values + 1e-10)
        if prior_rule != 1 and prior_weight > 0:
            prob = ((1 + score * prior_weight).unsqueeze(1) * recon).softmax(dim=1)
            prob = prob.log().clamp(CLAMP_LO, 0)
        else:
            prob = recon
        out_idx = log_sample_categorical(prob, uniform, return_index=True)
        out2 = x_idx.clone()
        pick = score.clone()
        if pick.sum() < 1e-6:
            pick += 1
        pick[x_idx != C - 1] = 0
        for i in range(B):
            n_sample = min(to_sample - sampled[i], prior_ps)
            if to_sample - sampled[i] - n_sample == 1:
                n_sample = to_sample - sampled[i]
            if n_sample <= 0:
                continue
            sel = multinomial_without_replacement(pick[i], n_sample, expo[i])
            out2[i][sel] = out_idx[i][sel]
            sampled[i] += int((out2[i] != C - 1).sum() - (x_idx[i] != C - 1).sum())
        return out2, sampled
    post = q_posterior(sched, recon, log_x_t, t)
    return log_sample_categorical(post, uniform, return_index=True), [1024] * B

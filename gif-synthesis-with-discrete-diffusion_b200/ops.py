"""Tensor-level wrappers of the C ABI: torch tensors in, torch tensors out, CUDA only.

Layout vocabulary (see include/d3pm_b200.h): the reference's logical `[B, C, N]` tensors (class
dim = 1) are held as *token-major rows* `[B, N, pitch]` with the class index contiguous.
`as_logical` / `rows_of` convert between the two without copying whenever the strides allow it.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from d3pm_b200 import _lib
from d3pm_b200._lib import D3PMError

LOGITS_DTYPES = {torch.float32: _lib.LOGITS_F32, torch.float16: _lib.LOGITS_F16, torch.bfloat16: _lib.LOGITS_BF16}
STREAM_CODEBOOKS = (1024, 2048, 4096)  # codebook sizes the persistent stream kernel is instantiated for
LOG_TINY = -69.07755278982137  # log(1e-30), the reference's one-hot "zero" (diffusion_transformer.py:50)


# --------------------------------------------------------------------------- plumbing
def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for x in tensors:
        if x is None:
            continue
        if not x.is_cuda:
            raise D3PMError("d3pm_b200 operates on CUDA tensors only (there is no CPU path)")
        if dev is not None and x.device != dev:
            raise D3PMError(f"tensors on different devices: {dev} vs {x.device}")
        dev = x.device
    if dev is None:
        raise D3PMError("no tensor given")
    return dev


def _ptr(x: Optional[torch.Tensor]) -> Optional[int]:
    return None if x is None else x.data_ptr()


def padded_pitch(num_classes: int) -> int:
    """Row pitch (floats) for a row of `num_classes` entries: next multiple of 4."""
    return (num_classes + 3) // 4 * 4


def alloc_rows(B: int, N: int, num_classes: int, device, dtype=torch.float32) -> torch.Tensor:
    """Uninitialised token-major storage `[B, N, pitch]`, rows 16-byte aligned."""
    return torch.empty(B, N, padded_pitch(num_classes), device=device, dtype=dtype)


def as_logical(rows: torch.Tensor, num_classes: int) -> torch.Tensor:
    """`[B, N, pitch]` rows -> the reference's logical `[B, C, N]` view (no copy)."""
    B, N, pitch = rows.shape
    return torch.as_strided(rows, (B, num_classes, N), (rows.stride(0), 1, rows.stride(1)), rows.storage_offset())


def rows_of(x: torch.Tensor) -> Optional[Tuple[torch.Tensor, int]]:
    """If logical `[B, C, N]` tensor `x` is a view of token-major rows, return `(rows [B,N,C], pitch)`."""
    if x.dim() != 3 or x.dtype not in LOGITS_DTYPES:
        return None
    B, C, N = x.shape
    sb, sc, sn = x.stride()
    if sc != 1 or (N > 1 and sn < C) or (B > 1 and sb != N * sn):
        return None
    return x.permute(0, 2, 1), (sn if N > 1 else C)


def to_rows(x: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """Logical `[B, C, N]` float32 -> token-major `(rows, pitch)`; copies (transposes on device) only
    when `x` is stored class-major like the reference's own `[B, K+1, N]` intermediates."""
    got = rows_of(x)
    if got is not None:
        return got
    dev = _need_cuda(x)
    B, C, N = x.shape
    src = x.contiguous().float()
    dst = alloc_rows(B, N, C, dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_to_token_major(src.data_ptr(), dst.data_ptr(), dst.shape[2], B, C, N, _stream(dev)),
               "d3pm_to_token_major")
    return dst[:, :, :C], dst.shape[2]


def new_status(device) -> torch.Tensor:
    return torch.zeros(1, dtype=torch.int32, device=device)


# --------------------------------------------------------------------------- schedule
def build_coef_table(sched8: torch.Tensor, T: int, K: int) -> torch.Tensor:
    """`[8, T+1]` float32 device schedule -> `[T, 32]` coefficient table (d3pm_build_coef_table)."""
    dev = _need_cuda(sched8)
    if sched8.shape != (8, T + 1) or sched8.dtype != torch.float32 or not sched8.is_contiguous():
        raise D3PMError(f"schedule must be a contiguous float32 [8, {T + 1}] tensor, got {tuple(sched8.shape)}")
    table = torch.empty(T, _lib.COEF_STRIDE, dtype=torch.float32, device=dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_build_coef_table(sched8.data_ptr(), T, K, table.data_ptr(), _stream(dev)),
               "d3pm_build_coef_table")
    return table


# --------------------------------------------------------------------------- fused step
def fused_step(logits_c: torch.Tensor, logits_u: Optional[torch.Tensor], x_t: torch.Tensor, t: torch.Tensor,
               coef_table: torch.Tensor, *, guidance_scale: float, sample_mode: int,
               gumbel: Optional[torch.Tensor] = None, gumbel_is_uniform: bool = False, seed: int = 0, offset: int = 0, row_offset: int = 0,
               want_post: bool = False, want_recon: bool = False, want_gap: bool = False,
               status: Optional[torch.Tensor] = None, thin_factor: float = 0.0,
               x_prev_out: Optional[torch.Tensor] = None, kernel: int = 0, sample_from: int = 0,
               want_score: bool = False, sharpen: Optional[torch.Tensor] = None,
               want_winner_post: bool = False) -> Dict[str, torch.Tensor]:
    """One fused reverse step over token-major logits `[B, N, K]` (d3pm_fused_step).

    `logits_u=None` is guidance off.  `gumbel` is `[B, N, >=K+1]` rows (entry K = [MASK]).
    Returns a dict with the requested tensors: `x_prev` int64 `[B, N]`, `post` / `recon` as
    `[B, N, pitch]` rows (use `as_logical(rows, K+1)` for the reference's `[B, K+1, N]`), `gap` `[B, N]`.
    Purity-prior sampling (p_sample with prior_rule 1 / 2): `sample_from=_lib.FROM_RECON` draws from p(x0 | x_t),
    `want_score` returns the per-token purity max_k p(x0 = k | x_t) as `score` `[B, N]`, `sharpen` `[B, N]` is the
    factor f of softmax(f * log p(x0 | x_t)).
    """
    dev = _need_cuda(logits_c, logits_u, x_t, t, coef_table, gumbel, status)
    if logits_c.dim() != 3 or logits_c.dtype not in LOGITS_DTYPES or logits_c.stride(2) != 1:
        raise D3PMError("logits_c must be float32 (or, for the production Philox step, float16 / bfloat16) [B, N, K] rows with the "
                        "class index contiguous")
    B, N, K = logits_c.shape
    pitch = logits_c.stride(1) if N > 1 else K
    if B > 1 and logits_c.stride(0) != N * pitch:
        raise D3PMError("logits_c rows must be uniformly pitched across the batch")
    if logits_u is not None and (logits_u.shape != logits_c.shape or logits_u.stride() != logits_c.stride()
                                 or logits_u.dtype != logits_c.dtype):
        raise D3PMError("logits_u must match logits_c in shape, strides and dtype")
    if x_t.shape != (B, N) or x_t.dtype != torch.int64 or not x_t.is_contiguous():
        raise D3PMError("x_t must be a contiguous int64 [B, N] tensor")
    if t.shape != (B,) or t.dtype != torch.int64 or not t.is_contiguous():
        raise D3PMError("t must be a contiguous int64 [B] tensor")
    T = coef_table.shape[0]

    d = _lib.StepDesc()
    d.logits_c, d.logits_u = _ptr(logits_c), _ptr(logits_u)
    d.x_t, d.t, d.coef_table = _ptr(x_t), _ptr(t), _ptr(coef_table)
    out: Dict[str, torch.Tensor] = {}
    if sample_mode != _lib.SAMPLE_NONE:
        if x_prev_out is None:
            x_prev_out = torch.empty(B, N, dtype=torch.int64, device=dev)
        elif x_prev_out.shape != (B, N) or x_prev_out.dtype != torch.int64 or not x_prev_out.is_contiguous():
            raise D3PMError("x_prev_out must be a contiguous int64 [B, N] tensor")
        out["x_prev"] = x_prev_out
        d.x_prev = _ptr(x_prev_out)
    if sample_mode == _lib.SAMPLE_GUMBEL:
        if gumbel is None or gumbel.dim() != 3 or gumbel.shape[:2] != (B, N) or gumbel.stride(2) != 1 \
                or gumbel.dtype != torch.float32:
            raise D3PMError("SAMPLE_GUMBEL needs float32 gumbel rows [B, N, >=K+1]")
        d.gumbel = _ptr(gumbel)
        d.pitch_gumbel = gumbel.stride(1) if N > 1 else gumbel.shape[2]
        d.gumbel_is_uniform = 1 if gumbel_is_uniform else 0
    pitch_out = padded_pitch(K + 1)
    if want_post:
        out["post"] = torch.empty(B, N, pitch_out, dtype=torch.float32, device=dev)
        d.post = _ptr(out["post"])
    if want_recon:
        out["recon"] = torch.empty(B, N, pitch_out, dtype=torch.float32, device=dev)
        d.recon = _ptr(out["recon"])
    if want_gap:
        out["gap"] = torch.empty(B, N, dtype=torch.float32, device=dev)
        d.gap = _ptr(out["gap"])
    if want_score:
        out["score"] = torch.empty(B, N, dtype=torch.float32, device=dev)
        d.score = _ptr(out["score"])
    if want_winner_post:  # stream kernel only: the log-posterior of the sampled class as that kernel computed it
        out["winner_post"] = torch.empty(B, N, dtype=torch.float32, device=dev)
        d.winner_post = _ptr(out["winner_post"])
    if sharpen is not None:
        if sharpen.shape != (B, N) or sharpen.dtype != torch.float32 or not sharpen.is_contiguous() or sharpen.device != dev:
            raise D3PMError("sharpen must be a contiguous float32 [B, N] tensor on the logits' device")
        d.sharpen = _ptr(sharpen)
    d.sample_from = int(sample_from)
    d.logits_dtype = LOGITS_DTYPES[logits_c.dtype]
    d.status = _ptr(status)
    d.B, d.N, d.K, d.T = B, N, K, T
    d.pitch_logits, d.pitch_out = pitch, pitch_out
    d.guidance_scale, d.sample_mode = float(guidance_scale), int(sample_mode)
    d.seed, d.offset, d.row_offset = seed & (2**64 - 1), offset & (2**64 - 1), int(row_offset)
    d.thin_factor = float(thin_factor)
    d.kernel = int(kernel)
    d.stream = _stream(dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_fused_step(ctypes.byref(d)), "d3pm_fused_step")
    return out


def philox_uniform(B: int, N: int, K: int, *, seed: int, offset: int, row_offset: int = 0,
                   device="cuda") -> torch.Tensor:
    """The uniforms SAMPLE_PHILOX draws, as rows `[B, N, pitch]` (first K+1 valid)."""
    dev = torch.device(device)
    u = alloc_rows(B, N, K + 1, dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_philox_uniform(u.data_ptr(), B * N, K, u.shape[2], seed & (2**64 - 1),
                                       offset & (2**64 - 1), row_offset, _stream(dev)), "d3pm_philox_uniform")
    return u


# --------------------------------------------------------------------------- fine-grained operators
def q_posterior_rows(log_x_start_rows: torch.Tensor, pitch_in: int, x_t: torch.Tensor, t: torch.Tensor,
                     coef_table: torch.Tensor, K: int, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q_posterior on arbitrary log p(x0) rows `[B, N, >=K]`; returns posterior rows `[B, N, pitch]`."""
    dev = _need_cuda(log_x_start_rows, x_t, t, coef_table, status)
    B, N = x_t.shape
    post = alloc_rows(B, N, K + 1, dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_q_posterior(log_x_start_rows.data_ptr(), pitch_in, x_t.data_ptr(), t.data_ptr(),
                                    coef_table.data_ptr(), post.data_ptr(), post.shape[2], B, N, K,
                                    coef_table.shape[0], _ptr(status), _stream(dev)), "d3pm_q_posterior")
    return post


def gumbel_argmax_rows(logits_rows: torch.Tensor, pitch_logits: int, num_classes: int, *,
                       noise_rows: Optional[torch.Tensor] = None, pitch_noise: int = 0, noise_kind: int = 2,
                       seed: int = 0, offset: int = 0, row_offset: int = 0, want_gap: bool = False):
    """log_sample_categorical over rows: `noise_kind` 0 = Gumbel given, 1 = uniform given, 2 = Philox."""
    dev = _need_cuda(logits_rows, noise_rows)
    B, N = logits_rows.shape[:2]
    x = torch.empty(B, N, dtype=torch.int64, device=dev)
    gap = torch.empty(B, N, dtype=torch.float32, device=dev) if want_gap else None
    lib = _lib.load_library()
    _lib.check(lib.d3pm_gumbel_argmax(logits_rows.data_ptr(), pitch_logits, _ptr(noise_rows), pitch_noise,
                                      noise_kind, x.data_ptr(), _ptr(gap), B * N, num_classes,
                                      seed & (2**64 - 1), offset & (2**64 - 1), row_offset, _stream(dev)),
               "d3pm_gumbel_argmax")
    return (x, gap) if want_gap else x


def purity_select(x_t: torch.Tensor, x_cand: torch.Tensor, score: Optional[torch.Tensor], n_reveal: torch.Tensor, K: int,
                  *, expo: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0, row_offset: int = 0):
    """The per-video reveal of the purity-prior branch (d3pm_purity_select): -> (x_out int64 `[B, N]`,
    revealed int32 `[B]`).  `score=None` is prior_rule 1 (uniform weights over the [MASK] positions); `expo` injects
    the Exp(1) noise of `torch.multinomial` (parity tests), otherwise it comes from the Philox stream."""
    dev = _need_cuda(x_t, x_cand, score, n_reveal, expo)
    B, N = x_t.shape
    for name, ten, dt in (("x_t", x_t, torch.int64), ("x_cand", x_cand, torch.int64), ("score", score, torch.float32),
                          ("expo", expo, torch.float32)):
        if ten is not None and (ten.shape != (B, N) or ten.dtype != dt or not ten.is_contiguous()):
            raise D3PMError(f"{name} must be a contiguous {dt} [B, N] tensor")
    if n_reveal.shape != (B,) or n_reveal.dtype != torch.int32 or not n_reveal.is_contiguous():
        raise D3PMError("n_reveal must be a contiguous int32 [B] tensor")
    x_out = torch.empty_like(x_t)
    revealed = torch.empty(B, dtype=torch.int32, device=dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_purity_select(x_t.data_ptr(), x_cand.data_ptr(), _ptr(score), _ptr(expo), n_reveal.data_ptr(),
                                      x_out.data_ptr(), revealed.data_ptr(), B, N, K, seed & (2**64 - 1),
                                      offset & (2**64 - 1), int(row_offset), _stream(dev)), "d3pm_purity_select")
    return x_out, revealed


def tokens_to_log_onehot_rows(x: torch.Tensor, num_classes: int, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """index_to_log_onehot: int64 `[B, N]` -> rows `[B, N, pitch]` holding {0, log 1e-30}."""
    dev = _need_cuda(x, status)
    if x.dtype != torch.int64 or x.dim() != 2 or not x.is_contiguous():
        raise D3PMError("tokens must be a contiguous int64 [B, N] tensor")
    B, N = x.shape
    out = alloc_rows(B, N, num_classes, dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_tokens_to_log_onehot(x.data_ptr(), out.data_ptr(), out.shape[2], B * N, num_classes,
                                             _ptr(status), _stream(dev)), "d3pm_tokens_to_log_onehot")
    return out


def argmax_classes(x: torch.Tensor) -> torch.Tensor:
    """log_onehot_to_index for any strided float32 `[B, C, N]` view -> int64 `[B, N]`."""
    dev = _need_cuda(x)
    if x.dim() != 3 or x.dtype != torch.float32:
        raise D3PMError("argmax_classes expects a float32 [B, C, N] tensor")
    B, C, N = x.shape
    idx = torch.empty(B, N, dtype=torch.int64, device=dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_argmax_classes(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), idx.data_ptr(),
                                       B, C, N, _stream(dev)), "d3pm_argmax_classes")
    return idx


# --------------------------------------------------------------------------- host-buffer entry (end-to-end path)
class HostStep:
    """The fused step for callers whose logits live in HOST memory (the reference's CPU tensors).

    A thin wrapper of the C handle `d3pm_host_step` (include/d3pm_b200.h): the library owns the device staging buffers
    for one batch shape, a copy stream and a compute stream; `__call__` hands it the HOST pointers, the inputs travel
    in chunks of whole videos and each chunk's fused step (production Philox sampling) overlaps the next transfer, the
    int64 tokens come back into a pinned host tensor and the call returns when they are there.  Per call it moves
    `h2d_bytes` up and `d2h_bytes` down; this is the path `bench.py` reports as `e2e`.

    `hidden_dim=64` makes it the host entry of the fused head instead (`d3pm_host_head_step_run`): the inputs are the
    hidden states `[B, N, 64]` that enter `to_logits`, 64x fewer bytes on the bus than the logits.
    """

    def __init__(self, B: int, N: int, K: int, coef_table: torch.Tensor, guidance: bool = True, *, T: Optional[int] = None,
                 hidden_dim: int = 0, chunks: int = 0, logits_dtype: torch.dtype = torch.float32):
        dev = coef_table.device
        if dev.type != "cuda":
            raise D3PMError("HostStep needs the coefficient table on a CUDA device")
        self.coef_table, self.guidance, self.shape, self.hidden_dim = coef_table, guidance, (B, N, K), hidden_dim
        self.device = dev
        self._lib = _lib.load_library()
        handle = ctypes.c_void_p()
        _lib.check(self._lib.d3pm_host_step_create(ctypes.byref(handle), dev.index if dev.index is not None else torch.cuda.current_device(),
                                                   B, N, K, int(T if T is not None else coef_table.shape[0]), hidden_dim,
                                                   1 if guidance else 0, chunks), "d3pm_host_step_create")
        self._handle = handle
        self.logits_dtype = logits_dtype
        if logits_dtype != torch.float32:  # host logits in float16 / bfloat16: half the bytes on the bus, stepped in place
            if hidden_dim or logits_dtype not in LOGITS_DTYPES:
                raise D3PMError("logits_dtype applies to a logits handle (hidden_dim = 0) and must be float32, float16 or bfloat16")
            _lib.check(self._lib.d3pm_host_step_set_logits_dtype(handle, LOGITS_DTYPES[logits_dtype]), "d3pm_host_step_set_logits_dtype")
        self.x_prev_host = torch.empty(B, N, dtype=torch.int64).pin_memory()
        self.h2d_bytes = int(self._lib.d3pm_host_step_h2d_bytes(handle))
        self.d2h_bytes = int(self._lib.d3pm_host_step_d2h_bytes(handle))
        self.last_status = 0

    def close(self):
        if getattr(self, "_handle", None) is not None:
            self._lib.d3pm_host_step_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_host(self, width, a, b, x_t, t):
        B, N, _ = self.shape
        fdt = torch.float32 if self.hidden_dim else self.logits_dtype
        for name, ten, shape, dt in (("conditional input", a, (B, N, width), fdt), ("unconditional input", b, (B, N, width), fdt),
                                     ("x_t", x_t, (B, N), torch.int64), ("t", t, (B,), torch.int64)):
            if ten is None:
                continue
            if ten.is_cuda:
                raise D3PMError("HostStep takes host tensors; use fused_step for device-resident inputs")
            if tuple(ten.shape) != shape or ten.dtype != dt or not ten.is_contiguous():
                raise D3PMError(f"{name} must be a contiguous {dt} tensor of shape {shape}")
        if self.guidance != (b is not None):
            raise D3PMError("the unconditional input must be given exactly when the handle was created with guidance")

    def __call__(self, logits_c: torch.Tensor, logits_u: Optional[torch.Tensor], x_t: torch.Tensor, t: torch.Tensor,
                 *, guidance_scale: float, seed: int, offset: int, row_offset: int = 0) -> torch.Tensor:
        if self.hidden_dim:
            raise D3PMError("this handle stages hidden states; call .head(...)")
        self._check_host(self.shape[2], logits_c, logits_u, x_t, t)
        st = ctypes.c_uint32(0)
        _lib.check(self._lib.d3pm_host_step_run(self._handle, logits_c.data_ptr(), _ptr(logits_u), x_t.data_ptr(), t.data_ptr(),
                                                self.coef_table.data_ptr(), float(guidance_scale), seed & (2**64 - 1),
                                                offset & (2**64 - 1), int(row_offset), self.x_prev_host.data_ptr(),
                                                ctypes.byref(st)), "d3pm_host_step_run")
        self.last_status = int(st.value)
        return self.x_prev_host

    def head(self, hw, hidden_c: torch.Tensor, hidden_u: Optional[torch.Tensor], x_t: torch.Tensor, t: torch.Tensor,
             *, guidance_scale: float, seed: int, offset: int, row_offset: int = 0) -> torch.Tensor:
        """Host hidden states -> tokens through the fused head (`hw`: `head.HeadWeights` on this device)."""
        if not self.hidden_dim:
            raise D3PMError("this handle stages logits; call it directly")
        self._check_host(self.hidden_dim, hidden_c, hidden_u, x_t, t)
        st = ctypes.c_uint32(0)
        _lib.check(self._lib.d3pm_host_head_step_run(self._handle, hidden_c.data_ptr(), _ptr(hidden_u), x_t.data_ptr(), t.data_ptr(),
                                                     hw.ln_weight.data_ptr(), hw.ln_bias.data_ptr(), float(hw.ln_eps),
                                                     hw.w_image.data_ptr(), hw.bias2.data_ptr(), self.coef_table.data_ptr(),
                                                     float(guidance_scale), float(hw.stat_slack(guidance_scale if hidden_u is not None else None)),
                                                     seed & (2**64 - 1), offset & (2**64 - 1),
                                                     int(row_offset), self.x_prev_host.data_ptr(), ctypes.byref(st)),
                   "d3pm_host_head_step_run")
        self.last_status = int(st.value)
        return self.x_prev_host

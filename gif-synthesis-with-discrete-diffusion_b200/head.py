"""Fused denoiser head + reverse step: tensor-level wrappers of `d3pm_head_prepare` / `d3pm_head_step`.

The reference's denoiser ends in `to_logits = Sequential(LayerNorm(n_embd), Linear(n_embd, K))`
(transformer_utils.py:352-356, applied at :441).  `HeadWeights` turns that module's parameters into the layout the
tensor-core kernel reads; `head_step` then takes the hidden states that FEED the head (`[B, N, n_embd]`, conditional and
unconditional pass) and returns x_{t-1} directly: the `[B, N, K]` logits (2 x 1 GiB per step at the benchmark shape)
are never written to memory.  CUDA only; nothing here falls back to PyTorch math.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import torch

from d3pm_b200 import _lib, ops
from d3pm_b200._lib import D3PMError


class HeadWeights:
    """Device image of `to_logits` for the fused kernel (built once per weight version).

    `valid` tells whether the fused kernel may be used: it evaluates classifier-free guidance in front of the Linear,
    which is exact only when no log-softmax entry can reach the reference's -70 clamp (diffusion_transformer.py:236),
    i.e. when every |logit| <= (70 - ln K) / 2.  The bound |logit| <= max_k ||W_k|| (sqrt(D) max|gamma| + ||beta||) +
    max|b| holds for every input because LayerNorm's normalised vector has norm <= sqrt(D)."""

    def __init__(self, ln_weight: torch.Tensor, ln_bias: torch.Tensor, ln_eps: float, weight: torch.Tensor,
                 bias: Optional[torch.Tensor]):
        dev = ops._need_cuda(ln_weight, ln_bias, weight, bias)
        K, D = weight.shape
        lib = _lib.load_library()
        n = int(lib.d3pm_head_image_floats(K, D))
        if n <= 0:
            raise D3PMError(f"the fused head supports n_embd = 64 and K in (1024, 2048, 4096); got D={D}, K={K}")
        self.K, self.D, self.ln_eps = K, D, float(ln_eps)
        self.ln_weight = ln_weight.detach().float().contiguous()
        self.ln_bias = ln_bias.detach().float().contiguous()
        self.w_image = torch.empty(n, dtype=torch.float32, device=dev)
        self.bias2 = torch.empty(K, dtype=torch.float32, device=dev)
        stats = torch.empty(2, dtype=torch.float32, device=dev)
        w = weight.detach().float().contiguous()
        b = None if bias is None else bias.detach().float().contiguous()
        _lib.check(lib.d3pm_head_prepare(w.data_ptr(), ops._ptr(b), K, D, self.w_image.data_ptr(), self.bias2.data_ptr(),
                                         stats.data_ptr(), ops._stream(dev)), "d3pm_head_prepare")
        max_norm, max_bias = (float(v) for v in stats.tolist())  # one host sync per weight version
        a_norm = math.sqrt(D) * float(self.ln_weight.abs().max()) + float(self.ln_bias.norm())
        self.logit_bound = max_norm * a_norm + max_bias
        self.valid = self.logit_bound <= 0.5 * (69.99 - math.log(K))
        self._wa_log2 = max_norm * a_norm * math.log2(math.e)  # bound of |W_k . a| in log2 units for one LayerNorm output

    def stat_slack(self, guidance_scale: Optional[float]) -> float:
        """Bound, in log2 units, of |logit_1xTF32 - logit| for the statistics pass of the fused kernel: the hi parts drop
        < 2^-10 of each factor, so the product sum is off by at most 2^-9 ||a|| ||W_k|| (Cauchy-Schwarz), with ||a|| the
        norm of the guidance-combined LayerNorm output, <= (|s| + |1 - s|) times that of one output."""
        mix = 1.0 if guidance_scale is None else abs(guidance_scale) + abs(1.0 - guidance_scale)
        return 1.05 * mix * self._wa_log2 / 512.0 + 1e-3

    @classmethod
    def from_module(cls, to_logits: torch.nn.Module) -> "HeadWeights":
        ln, lin = to_logits[0], to_logits[-1]
        if not isinstance(ln, torch.nn.LayerNorm) or not isinstance(lin, torch.nn.Linear) or len(to_logits) != 2:
            raise D3PMError("to_logits must be Sequential(LayerNorm, Linear) (transformer_utils.py:352-356)")
        if ln.weight is None or ln.bias is None:
            raise D3PMError("to_logits[0] must be an affine LayerNorm")
        return cls(ln.weight, ln.bias, ln.eps, lin.weight, lin.bias)


def head_step(hw: HeadWeights, hidden_c: torch.Tensor, hidden_u: Optional[torch.Tensor], x_t: Optional[torch.Tensor],
              t: Optional[torch.Tensor], coef_table: Optional[torch.Tensor], *, guidance_scale: float, mode: int = _lib.HEAD_STEP,
              seed: int = 0, offset: int = 0, row_offset: int = 0, status: Optional[torch.Tensor] = None,
              x_prev_out: Optional[torch.Tensor] = None, thin_factor: float = 0.0, scratch=None, stats_1xtf32: bool = True):
    """One reverse step from the hidden states `[B, N, D]` that feed `to_logits` (d3pm_head_step).

    mode HEAD_STEP / HEAD_REFERENCE -> int64 tokens `[B, N]`; HEAD_LOGITS -> the guidance-combined logits `[B, N, K]`.
    `scratch` (from `head_scratch`) avoids per-call allocations in a loop."""
    dev = ops._need_cuda(hidden_c, hidden_u, x_t, t, coef_table, status)
    if hidden_c.dim() != 3 or hidden_c.dtype != torch.float32 or not hidden_c.is_contiguous() or hidden_c.shape[2] != hw.D:
        raise D3PMError(f"hidden_c must be a contiguous float32 [B, N, {hw.D}] tensor")
    if hidden_u is not None and (hidden_u.shape != hidden_c.shape or hidden_u.dtype != torch.float32 or not hidden_u.is_contiguous()):
        raise D3PMError("hidden_u must match hidden_c")
    B, N, D = hidden_c.shape
    d = _lib.HeadDesc()
    d.hidden_c, d.hidden_u = hidden_c.data_ptr(), ops._ptr(hidden_u)
    d.ln_weight, d.ln_bias, d.w_image, d.bias2 = hw.ln_weight.data_ptr(), hw.ln_bias.data_ptr(), hw.w_image.data_ptr(), hw.bias2.data_ptr()
    d.B, d.N, d.K, d.D, d.mode = B, N, hw.K, D, int(mode)
    d.ln_eps, d.guidance_scale, d.thin_factor = hw.ln_eps, float(guidance_scale), float(thin_factor)
    # statistics pass in 1xTF32 (a third of its tensor work); `stats_1xtf32=False` keeps it in 3xTF32 (same tokens, tested)
    d.stat_slack = hw.stat_slack(guidance_scale if hidden_u is not None else None) if stats_1xtf32 else 0.0
    d.seed, d.offset, d.row_offset = seed & (2**64 - 1), offset & (2**64 - 1), int(row_offset)
    d.status, d.stream = ops._ptr(status), ops._stream(dev)
    if mode == _lib.HEAD_LOGITS:
        out = torch.empty(B, N, hw.K, dtype=torch.float32, device=dev)
        d.logits_out = out.data_ptr()
    else:
        if x_t is None or t is None or coef_table is None:
            raise D3PMError("x_t, t and coef_table are required")
        if x_t.shape != (B, N) or x_t.dtype != torch.int64 or not x_t.is_contiguous():
            raise D3PMError("x_t must be a contiguous int64 [B, N] tensor")
        if t.shape != (B,) or t.dtype != torch.int64 or not t.is_contiguous():
            raise D3PMError("t must be a contiguous int64 [B] tensor")
        out = x_prev_out if x_prev_out is not None else torch.empty(B, N, dtype=torch.int64, device=dev)
        if out.shape != (B, N) or out.dtype != torch.int64 or not out.is_contiguous():
            raise D3PMError("x_prev_out must be a contiguous int64 [B, N] tensor")
        d.x_t, d.t, d.coef_table, d.T, d.x_prev = x_t.data_ptr(), t.data_ptr(), coef_table.data_ptr(), coef_table.shape[0], out.data_ptr()
        if scratch is None:
            scratch = head_scratch(B, N, dev)
        d.redo_rows, d.redo_count = scratch[0].data_ptr(), scratch[1].data_ptr()
    lib = _lib.load_library()
    _lib.check(lib.d3pm_head_step(ctypes.byref(d)), "d3pm_head_step")
    return out


def head_scratch(B: int, N: int, device):
    return (torch.empty(B * N, dtype=torch.int32, device=device), torch.zeros(1, dtype=torch.int32, device=device))

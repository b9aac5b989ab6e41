"""Token -> video, first stage: the consumer of the int64 `[B, N]` tokens this path samples.

`VQVAE.decode` (videogpt_vq_vae.py:53-56) is `decoder(post_vq_conv(shift_dim(F.embedding(tokens, codebook), -1, 1)))`.
The embedding gather and the 1x1x1 convolution are both per-token, so they fold into one `[K, C]` table and a single
gather kernel that writes the channels-first tensor the decoder's convolutions take; the 3-D (transposed-)convolution
decoder itself stays the reference's PyTorch module.  CUDA only.
"""
from __future__ import annotations

from typing import Optional

import torch

from d3pm_b200 import _lib, ops
from d3pm_b200._lib import D3PMError


class DecodeTable:
    """`lut[k] = post_vq_conv(codebook[k])`, rebuilt when the codebook or the convolution changes."""

    def __init__(self, codebook: torch.Tensor, conv_weight: torch.Tensor, conv_bias: Optional[torch.Tensor]):
        dev = ops._need_cuda(codebook, conv_weight, conv_bias)
        K, E = codebook.shape
        w = conv_weight.detach().float().reshape(conv_weight.shape[0], -1).contiguous()
        if w.shape[1] != E:
            raise D3PMError(f"post_vq_conv must be a 1x1x1 convolution over the {E} embedding channels, got weight {tuple(conv_weight.shape)}")
        C = w.shape[0]
        cb = codebook.detach().float().contiguous()
        b = None if conv_bias is None else conv_bias.detach().float().contiguous()
        self.K, self.E, self.C = K, E, C
        self.lut = torch.empty(K, C, dtype=torch.float32, device=dev)
        lib = _lib.load_library()
        _lib.check(lib.d3pm_decode_lut(cb.data_ptr(), w.data_ptr(), ops._ptr(b), K, E, C, self.lut.data_ptr(), ops._stream(dev)),
                   "d3pm_decode_lut")

    @classmethod
    def from_autoencoder(cls, autoencoder) -> "DecodeTable":
        """`autoencoder` = the reference's `VQVAE`: `.codebook.embeddings` `[K, E]`, `.post_vq_conv.conv` (Conv3d, kernel 1)."""
        conv = autoencoder.post_vq_conv.conv
        if tuple(conv.kernel_size) != (1, 1, 1) or tuple(conv.stride) != (1, 1, 1):
            raise D3PMError("post_vq_conv must be a 1x1x1, stride-1 convolution (videogpt_vq_vae.py:31)")
        return cls(autoencoder.codebook.embeddings, conv.weight, conv.bias)


def tokens_to_features(table: DecodeTable, tokens: torch.Tensor, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int64 tokens `[B, T', H', W']` (or `[B, N]`) -> `post_vq_conv(embedding)` as float32 `[B, C, T', H', W']`."""
    dev = ops._need_cuda(tokens, table.lut, status)
    if tokens.dtype != torch.int64 or tokens.dim() < 2:
        raise D3PMError("tokens must be an int64 tensor [B, ...]")
    B = tokens.shape[0]
    flat = tokens.reshape(B, -1).contiguous()
    N = flat.shape[1]
    out = torch.empty(B, table.C, N, dtype=torch.float32, device=dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_tokens_to_features(flat.data_ptr(), table.lut.data_ptr(), out.data_ptr(), B, N, table.K, table.C,
                                           ops._ptr(status), ops._stream(dev)), "d3pm_tokens_to_features")
    return out.view(B, table.C, *tokens.shape[1:])


def decode(autoencoder, tokens: torch.Tensor, table: Optional[DecodeTable] = None) -> torch.Tensor:
    """Drop-in for `VQVAE.decode(encodings)` (videogpt_vq_vae.py:53-56): fused gather + 1x1x1 conv, then the reference's
    own `decoder` module."""
    table = table if table is not None else DecodeTable.from_autoencoder(autoencoder)
    return autoencoder.decoder(tokens_to_features(table, tokens))

"""CPU restatement of the reference's VQ-VAE `Decoder` in eval mode.  TEST INFRASTRUCTURE ONLY: imported by
`tests/` (and nothing in the product package); pinned to outputs of the imported reference by `tests/golden/decode_*.npz`
(`tests/test_oracle_decode_golden.py`).

Follows, op for op but without the reference's modules,
    src/models/networks/videogpt_vq_vae.py   Decoder.forward :281-287, AttentionResidualBlock :120-136, AxialBlock :100-118,
                                             SamePadConv3d :289-310, SamePadConvTranspose3d :312-334, VQVAE.decode :53-56
    src/models/utils/model_utils.py          MultiHeadAttention.forward :238-285, AxialAttention :318-336,
                                             scaled_dot_product_attention :586-600, shift_dim :17-39
on a `state_dict` of the reference's `Decoder` (keys `res_stack.{i}.block.{j}...`, `convts.{i}.convt...`).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm3d default


def same_pad(kernel: Sequence[int], stride: Sequence[int]):
    """`pad_input` of SamePadConv3d / SamePadConvTranspose3d (:296-301, :319-323): total k - s per dimension, the odd
    element in front, listed last dimension first as F.pad takes it."""
    pad = []
    for k, s in list(zip(kernel, stride))[::-1]:
        p = k - s
        pad += [p // 2 + p % 2, p // 2]
    return tuple(pad)


def batch_norm_eval(x, sd, prefix):
    """nn.BatchNorm3d in eval mode: running statistics, per channel (dim 1)."""
    shape = (1, -1, 1, 1, 1)
    mean, var = sd[prefix + "running_mean"].view(shape), sd[prefix + "running_var"].view(shape)
    w, b = sd[prefix + "weight"].view(shape), sd[prefix + "bias"].view(shape)
    return (x - mean) / torch.sqrt(var + BN_EPS) * w + b


def attention_along(x_last, sd, prefix, axis, n_head=2):
    """MultiHeadAttention with AxialAttention along `axis` (1 = t, 2 = h, 3 = w) of a channels-last `[b, t, h, w, c]`
    tensor: Linear q / k / v without bias, heads split off the channel dim, softmax(q k^T / sqrt(d)) v along the axis,
    heads merged, Linear `fc` with bias."""
    c = x_last.shape[-1]
    d = c // n_head
    q = F.linear(x_last, sd[prefix + "w_qs.weight"]).unflatten(-1, (n_head, d))
    k = F.linear(x_last, sd[prefix + "w_ks.weight"]).unflatten(-1, (n_head, d))
    v = F.linear(x_last, sd[prefix + "w_vs.weight"]).unflatten(-1, (n_head, d))
    # [b, t, h, w, head, d] -> the attended axis next to d
    q, k, v = (z.movedim(axis, -2) for z in (q, k, v))          # [..., head, L, d] with the other two grid dims in front
    attn = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    a = torch.matmul(F.softmax(attn, dim=-1), v)                  # [..., head, L, d]
    a = a.movedim(-2, axis)                                       # back to [b, t, h, w, head, d]
    return F.linear(a.flatten(-2), sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])


def residual_block(x, sd, prefix):
    """AttentionResidualBlock (:120-136): x + AxialBlock(ReLU(BN(conv1(ReLU(BN(conv3(ReLU(BN(x))))))))) ."""
    h = F.relu(batch_norm_eval(x, sd, prefix + "block.0."))
    h = F.conv3d(F.pad(h, same_pad((3, 3, 3), (1, 1, 1))), sd[prefix + "block.2.conv.weight"])
    h = F.relu(batch_norm_eval(h, sd, prefix + "block.3."))
    h = F.conv3d(h, sd[prefix + "block.5.conv.weight"])
    h = F.relu(batch_norm_eval(h, sd, prefix + "block.6."))
    hl = h.permute(0, 2, 3, 4, 1)  # shift_dim(x, 1, -1)
    a = (attention_along(hl, sd, prefix + "block.8.attn_w.", 3) + attention_along(hl, sd, prefix + "block.8.attn_h.", 2)
         + attention_along(hl, sd, prefix + "block.8.attn_t.", 1))
    return x + a.permute(0, 4, 1, 2, 3)


def decoder_forward(sd: Dict[str, torch.Tensor], h: torch.Tensor, n_res_layers: int, strides: Sequence[Sequence[int]]) -> torch.Tensor:
    """`Decoder.forward` (:281-287) on `h` `[B, C, T, H, W]`; `strides[i]` = stride of `convts[i]` (kernel 4)."""
    x = h
    for i in range(n_res_layers):
        x = residual_block(x, sd, f"res_stack.{i}.")
    x = F.relu(batch_norm_eval(x, sd, f"res_stack.{n_res_layers}."))
    for i, s in enumerate(strides):
        x = F.conv_transpose3d(F.pad(x, same_pad((4, 4, 4), s)), sd[f"convts.{i}.convt.weight"], sd[f"convts.{i}.convt.bias"],
                               stride=tuple(s), padding=(3, 3, 3))
        if i < len(strides) - 1:
            x = F.relu(x)
    return x


def upsample_strides(downsample: Sequence[int]):
    """The strides Decoder.__init__ (:268-278) gives its transposed convolutions for `downsample` (powers of two)."""
    n = [int(math.log2(d)) for d in downsample]
    out = []
    for _ in range(max(n)):
        out.append(tuple(2 if d > 0 else 1 for d in n))
        n = [d - 1 for d in n]
    return out


def vqvae_decode(sd_vq: Dict[str, torch.Tensor], tokens: torch.Tensor, n_res_layers: int, downsample: Sequence[int]) -> torch.Tensor:
    """`VQVAE.decode` (:53-56) on a `state_dict` of the reference's `VQVAE`."""
    h = F.embedding(tokens, sd_vq["codebook.embeddings"]).movedim(-1, 1)
    h = F.conv3d(h, sd_vq["post_vq_conv.conv.weight"], sd_vq["post_vq_conv.conv.bias"])
    dec = {k[len("decoder."):]: v for k, v in sd_vq.items() if k.startswith("decoder.")}
    return decoder_forward(dec, h, n_res_layers, upsample_strides(downsample))

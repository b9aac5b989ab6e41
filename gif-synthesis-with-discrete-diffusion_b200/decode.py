"""Token -> video: the consumer of the int64 `[B, N]` tokens this path samples.

`VQVAE.decode` (videogpt_vq_vae.py:53-56) is `decoder(post_vq_conv(shift_dim(F.embedding(tokens, codebook), -1, 1)))`.

* First stage: the embedding gather and the 1x1x1 convolution are both per-token, so they fold into one `[K, C]` table
  (`DecodeTable`) and a single gather kernel.
* Second stage: `NativeDecoder` runs the reference's `Decoder` (:258-287; eval mode) layer by layer on channels-last
  activations - every convolution, transposed convolution and Linear is `d3pm_dec_conv` (tcgen05 implicit GEMM, 3xTF32),
  the axial attentions `d3pm_dec_axial_attention`, the last 3-channel transposed convolution a GEMM + `d3pm_dec_col2im`.
  BatchNorm3d (eval) + ReLU pairs are folded: one that FOLLOWS a bias-free convolution into that convolution's weights and
  bias, one that PRECEDES a convolution into the per-channel affine the kernel applies while it gathers its A operand.

`decode(autoencoder, tokens, decoder=NativeDecoder...)` is the drop-in for `VQVAE.decode`; without `decoder` the reference's
own PyTorch `Decoder` module runs behind the fused first stage.  CUDA only.
"""
from __future__ import annotations

import ctypes
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


def embed_rows(table: DecodeTable, tokens: torch.Tensor, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int64 tokens `[B, ...]` -> channels-last rows `[B * N, C]` of `post_vq_conv(embedding)` (the native decoder's input)."""
    dev = ops._need_cuda(tokens, table.lut, status)
    if tokens.dtype != torch.int64 or tokens.dim() < 2:
        raise D3PMError("tokens must be an int64 tensor [B, ...]")
    flat = tokens.reshape(-1).contiguous()
    out = torch.empty(flat.numel(), table.C, dtype=torch.float32, device=dev)
    lib = _lib.load_library()
    _lib.check(lib.d3pm_dec_embed_rows(flat.data_ptr(), table.lut.data_ptr(), out.data_ptr(), flat.numel(), table.K, table.C,
                                       ops._ptr(status), ops._stream(dev)), "d3pm_dec_embed_rows")
    return out


def _bn_affine(bn) -> tuple:
    """Eval-mode BatchNorm3d as y = x * scale + shift (float64 arithmetic, fp32 result)."""
    var, mean = bn.running_var.detach().double(), bn.running_mean.detach().double()
    w = bn.weight.detach().double() if bn.weight is not None else torch.ones_like(var)
    b = bn.bias.detach().double() if bn.bias is not None else torch.zeros_like(var)
    scale = w / torch.sqrt(var + bn.eps)
    return scale, b - mean * scale


def _convt_taps(stride):
    """Parity classes of SamePadConvTranspose3d (kernel 4, padding as :319-332) with `stride` in {1, 2} per dimension.
    Along one dimension, output y = m * s + par receives x[m + d] * w[k] for every k with (par + 3 - k) % s == 0,
    d = (par + 3 - k) / s - ceil((4 - s) / 2).  Returns [(class offsets (pt, ph, pw), [((kt, kh, kw), (dt, dh, dw)), ...])]."""
    per_dim = []
    for s in stride:
        if s not in (1, 2):
            raise D3PMError(f"transposed-convolution stride {tuple(stride)}: only 1 and 2 are built")
        pf = (4 - s + 1) // 2
        per_dim.append([[(k, (par + 3 - k) // s - pf) for k in range(4) if (par + 3 - k) % s == 0] for par in range(s)])
    classes = []
    for pt in range(stride[0]):
        for ph in range(stride[1]):
            for pw in range(stride[2]):
                taps = [((kt, kh, kw), (dt, dh, dw)) for kt, dt in per_dim[0][pt] for kh, dh in per_dim[1][ph]
                        for kw, dw in per_dim[2][pw]]
                classes.append(((pt, ph, pw), taps))
    return classes


class LayerSpec:
    """One `d3pm_dec_conv` call, device-independent: the K-major weight matrix `[nclass, Nout, ntaps * cin]`, the tap
    offsets of every parity class, and what surrounds the product (input affine + ReLU, bias, output ReLU)."""

    def __init__(self, wmat, *, cin, taps, classes, stride=(1, 1, 1), bias=None, in_affine=None, relu_out=False):
        nclass, nout, ktot = wmat.shape
        assert ktot == len(taps[0]) * cin and nclass == len(classes) == len(taps)
        assert nclass <= _lib.DEC_MAX_CLASSES and len(taps[0]) <= _lib.DEC_MAX_TAPS
        self.wmat, self.cin, self.nout, self.taps, self.classes = wmat.float().contiguous(), cin, nout, taps, classes
        self.stride, self.relu_out = tuple(int(s) for s in stride), bool(relu_out)
        self.bias = None if bias is None else bias.float().contiguous()
        self.in_affine = None if in_affine is None else tuple(a.float().contiguous() for a in in_affine)


def decoder_plan(decoder: torch.nn.Module) -> dict:
    """Read the reference's `Decoder` module (eval mode) into the list of `LayerSpec`s `NativeDecoder` executes.  Pure torch,
    any device: the CPU tests run this plan through an emulation of the C contract against the oracle."""
    if decoder.training:
        raise D3PMError("NativeDecoder folds BatchNorm running statistics: put the decoder in eval() mode first "
                        "(train-mode batch statistics stay with the PyTorch module)")
    blocks = list(decoder.res_stack)
    bn_final, res_blocks = blocks[-2], blocks[:-2]
    C = bn_final.num_features
    if C % 2 != 0 or C > 256:
        raise D3PMError(f"n_hiddens={C}: the native decoder covers even n_hiddens up to 256 (two heads of up to 128 channels)")
    # The kernels work on multiples of 32 channels with heads of 32 / 64 / 128: every channel vector is ZERO-PADDED to
    # Cp = C rounded up to 64 (the half-width tensors and the two attention heads to Cp / 2).  Padded weights, biases and
    # BatchNorm affines are zero, so padded channels stay exactly zero through every layer and never reach a real one.
    Cp = (C + 63) // 64 * 64
    if Cp // 2 not in (32, 64, 128):
        Cp = 256 if Cp > 128 else 128   # (192 -> 256: heads of 96 become 128)
    Ch, Chp = C // 2, Cp // 2

    def pad(t, dim, n):
        if t.shape[dim] == n:
            return t
        shape = list(t.shape)
        shape[dim] = n - t.shape[dim]
        return torch.cat([t, torch.zeros(shape, dtype=t.dtype, device=t.device)], dim)

    def pad_heads(t, dim):   # a C-vector that is [2 heads][C / 2] -> [2][Cp / 2]
        shape = list(t.shape)
        t = t.reshape(shape[:dim] + [2, Ch] + shape[dim + 1:])
        t = pad(t, dim + 1, Chp)
        return t.reshape(shape[:dim] + [Cp] + shape[dim + 1:])

    def pad_affine(aff, n):
        return tuple(pad(a, 0, n) for a in aff)

    one = [(0, 0, 0)]
    conv3_taps = [(kt - 1, kh - 1, kw - 1) for kt in range(3) for kh in range(3) for kw in range(3)]
    plan = {"C": Cp, "C_model": C, "heads": 2, "softmax_scale": float(Ch) ** -0.5, "blocks": [], "convts": []}
    with torch.no_grad():
        for rb in res_blocks:
            bn1, _, conv3, bn2, _, conv1, bn3, _, axial = list(rb.block)
            w3, w1 = conv3.conv.weight.detach().double(), conv1.conv.weight.detach().double()
            if tuple(w3.shape[2:]) != (3, 3, 3) or tuple(w1.shape[2:]) != (1, 1, 1) or conv3.conv.bias is not None or conv1.conv.bias is not None:
                raise D3PMError("AttentionResidualBlock: expected a bias-free 3x3x3 and a bias-free 1x1x1 convolution (:124-131)")
            s2, b2 = _bn_affine(bn2)
            s3, b3 = _bn_affine(bn3)
            w3 = pad(pad(w3 * s2.view(-1, 1, 1, 1, 1), 0, Chp), 1, Cp)
            w1 = pad(pad(w1 * s3.view(-1, 1, 1, 1, 1), 0, Cp), 1, Chp)
            m3 = w3.permute(0, 2, 3, 4, 1).reshape(1, Chp, -1)   # [n][tap][cin]
            m1 = w1.reshape(1, Cp, -1)
            L3 = LayerSpec(m3, cin=Cp, taps=[conv3_taps], classes=[(0, 0, 0)], bias=pad(b2, 0, Chp), in_affine=pad_affine(_bn_affine(bn1), Cp),
                           relu_out=True)
            L1 = LayerSpec(m1, cin=Chp, taps=[one], classes=[(0, 0, 0)], bias=pad(b3, 0, Cp), relu_out=True)
            if axial.attn_w.n_head != 2 or axial.attn_w.d_k != Ch:
                raise D3PMError("AxialBlock: expected two heads of n_hiddens / 2 channels (:103-111)")
            att = (axial.attn_w, axial.attn_h, axial.attn_t)  # axes W, H, T = the kernel's axis 0, 1, 2
            wqkv = torch.cat([torch.cat([pad(pad_heads(w.weight.detach(), 0), 1, Cp) for w in (a.w_qs, a.w_ks, a.w_vs)], 0) for a in att], 0)
            Lq = LayerSpec(wqkv.unsqueeze(0), cin=Cp, taps=[one], classes=[(0, 0, 0)])
            wfc = torch.cat([pad(pad_heads(a.fc.weight.detach(), 1), 0, Cp) for a in att], 1)
            bfc = pad(sum(a.fc.bias.detach().double() for a in att), 0, Cp)
            Lf = LayerSpec(wfc.unsqueeze(0), cin=3 * Cp, taps=[one], classes=[(0, 0, 0)], bias=bfc)
            plan["blocks"].append((L3, L1, Lq, Lf))
        in_aff = pad_affine(_bn_affine(bn_final), Cp)
        n = len(decoder.convts)
        for i, ct in enumerate(decoder.convts):
            w = ct.convt.weight.detach()   # [Cin, Cout, 4, 4, 4]
            stride = tuple(int(v) for v in ct.convt.stride)
            if tuple(w.shape[2:]) != (4, 4, 4) or tuple(ct.convt.padding) != (3, 3, 3) or w.shape[0] != C:
                raise D3PMError("SamePadConvTranspose3d: expected n_hiddens input channels, kernel 4 and padding 3 (:330-332)")
            bias = ct.convt.bias.detach() if ct.convt.bias is not None else torch.zeros(w.shape[1], device=w.device)
            w = pad(w, 0, Cp)
            if i < n - 1:
                if w.shape[1] != C:
                    raise D3PMError("SamePadConvTranspose3d: an inner layer is expected to keep n_hiddens channels (:271-275)")
                w, bias = pad(w, 1, Cp), pad(bias, 0, Cp)
                classes = _convt_taps(stride)
                mats = [torch.stack([w[:, :, kt, kh, kw].t() for (kt, kh, kw), _ in taps], 1).reshape(Cp, -1)
                        for _, taps in classes]   # [Cout][tap][Cin]
                spec = LayerSpec(torch.stack(mats, 0), cin=Cp, taps=[[d for _, d in taps] for _, taps in classes],
                                 classes=[c for c, _ in classes], stride=stride, bias=bias, in_affine=in_aff, relu_out=True)
                plan["convts"].append(("conv", spec, stride))
            else:
                cout = w.shape[1]
                if cout > 4:
                    raise D3PMError(f"last transposed convolution has {cout} output channels; col2im is built for <= 4 (RGB)")
                m = w.permute(2, 3, 4, 1, 0).reshape(1, 64 * cout, Cp)   # [(kt, kh, kw, co)][cin]
                spec = LayerSpec(m, cin=Cp, taps=[one], classes=[(0, 0, 0)], in_affine=in_aff)
                plan["convts"].append(("col2im", spec, stride, bias.float().contiguous(), cout))
            in_aff = None  # later layers read an already activated tensor (the ReLU sits in the previous epilogue)
    return plan


class _Layer:
    """A `LayerSpec` bound to a device: the weight image and the static part of the launch descriptor.  (The descriptor is
    reused between calls: one `NativeDecoder` serves one stream / thread at a time, like an `nn.Module` with buffers.)"""

    def __init__(self, spec: LayerSpec, dev, n_tile: Optional[int] = None, cta_pair: Optional[bool] = False):
        nclass, nout, ktot = spec.wmat.shape
        self.cin, self.nout, self.nclass, self.ntaps, self.stride = spec.cin, nout, nclass, len(spec.taps[0]), spec.stride
        self.n_tile = n_tile if n_tile is not None else (128 if nout <= 128 else 256)
        self.cta_pair = cta_pair if cta_pair is None else bool(cta_pair)   # None: decided per call from the number of tiles
        self.sms = torch.cuda.get_device_properties(dev).multi_processor_count
        npad = (nout + self.n_tile - 1) // self.n_tile * self.n_tile
        lib = _lib.load_library()
        self.image = torch.empty(lib.d3pm_dec_image_floats(nclass, nout, ktot, self.n_tile), dtype=torch.float32, device=dev)
        w = spec.wmat.to(dev)
        _lib.check(lib.d3pm_dec_weight_image(w.data_ptr(), nclass, nout, ktot, self.n_tile, self.image.data_ptr(), ops._stream(dev)),
                   "d3pm_dec_weight_image")
        torch.cuda.current_stream(dev).synchronize()  # `w` may be a temporary
        self.bias = None
        if spec.bias is not None:
            self.bias = torch.zeros(npad, dtype=torch.float32, device=dev)
            self.bias[:nout] = spec.bias.to(dev)
        self.in_scale = self.in_shift = None
        if spec.in_affine is not None:
            self.in_scale, self.in_shift = (a.to(dev) for a in spec.in_affine)
        self.relu_out = spec.relu_out
        self.desc = _lib.DecConvDesc()
        for c, (offs, tl) in enumerate(zip(spec.classes, spec.taps)):
            for e in range(3):
                self.desc.cls[c][e] = offs[e]
            for i, d in enumerate(tl):
                for e in range(3):
                    self.desc.tap[c][i][e] = d[e]

    def __call__(self, x: torch.Tensor, B: int, grid, *, terms: int, residual: Optional[torch.Tensor] = None,
                 transposed: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-> rows `[B * T' * H' * W', Nout]`, or with `transposed` planes `[B * T', Nout, H' * W']` (what d3pm_dec_col2im reads)."""
        T, H, W = grid
        rows_out = B * T * H * W * self.stride[0] * self.stride[1] * self.stride[2]
        if x.shape != (B * T * H * W, self.cin) or not x.is_contiguous() or x.dtype != torch.float32:
            raise D3PMError(f"expected contiguous float32 rows [{B * T * H * W}, {self.cin}], got {tuple(x.shape)}")
        plane = H * W * self.stride[1] * self.stride[2]
        if out is None:
            out = torch.empty((rows_out // plane, self.nout, plane) if transposed else (rows_out, self.nout), dtype=torch.float32,
                              device=x.device)
        elif transposed or out.dim() != 2 or out.shape[0] != rows_out or out.shape[1] < self.nout or out.stride(1) != 1 or out.stride(0) % 4:
            raise D3PMError("out must be float32 rows [output positions, >= Nout] with a pitch that is a multiple of 4")
        d = self.desc
        d.x, d.in_scale, d.in_shift = x.data_ptr(), ops._ptr(self.in_scale), ops._ptr(self.in_shift)
        if residual is not None and (residual.shape != out.shape or residual.stride() != out.stride()):
            raise D3PMError("residual must be laid out like the output rows")
        d.w_image, d.bias, d.residual, d.out = self.image.data_ptr(), ops._ptr(self.bias), ops._ptr(residual), out.data_ptr()
        d.B, d.T, d.H, d.W, d.Cin = B, T, H, W, self.cin
        d.ntaps, d.nclass, d.Nout, d.out_transposed = self.ntaps, self.nclass, self.nout, int(transposed)
        d.ldo = out.shape[-1] if transposed else out.stride(0)
        d.stride_t, d.stride_h, d.stride_w = self.stride
        pair = self.cta_pair
        if pair is None:
            # pairs of CTAs (cta_group::2) pay off when every pair still gets at least two 256-position tiles: the transposed
            # convolutions, the qkv product; the 3x3x3 convolution of a small batch keeps single CTAs (more tiles than SMs matter more)
            npad = (self.nout + self.n_tile - 1) // self.n_tile
            pair_tiles = ((B * T * H * W + 255) // 256) * npad * self.nclass
            pair = pair_tiles >= 2 * (self.sms // 2)
        d.relu_out, d.terms, d.n_tile, d.cta_pair = int(self.relu_out), terms, self.n_tile, int(pair)
        d.stream = ops._stream(x.device)
        _lib.check(_lib.load_library().d3pm_dec_conv(ctypes.byref(d)), "d3pm_dec_conv")
        return out


class NativeDecoder:
    """The reference's `Decoder` (videogpt_vq_vae.py:258-287) in eval mode on the library's kernels.

    `precision="fp32"`: every product as 3xTF32 on the tensor cores (fp32-grade: what the reference computes on the CPU, or
    on a GPU with TF32 disabled); `"tf32"`: single TF32 products, the accuracy class of the reference's default GPU path
    (cuDNN convolutions run TF32 unless `torch.backends.cudnn.allow_tf32 = False`), about three times less tensor work.
    Rebuild the object when the module's weights or BatchNorm statistics change (they are folded at construction).
    `cta_pair`: run the products on pairs of CTAs (tcgen05 cta_group::2) - None (default) decides per layer and batch, True /
    False force it; the results do not depend on it.
    """

    def __init__(self, decoder: torch.nn.Module, precision: str = "fp32", n_tile: Optional[int] = None, cta_pair: Optional[bool] = None):
        if precision not in ("fp32", "tf32"):
            raise D3PMError("precision must be 'fp32' (3xTF32) or 'tf32'")
        self.terms = 3 if precision == "fp32" else 1
        plan = decoder_plan(decoder)
        dev = next(decoder.parameters()).device
        if dev.type != "cuda":
            raise D3PMError("d3pm_b200 operates on CUDA tensors only (there is no CPU path)")
        self.C, self.C_model, self.heads, self.device = plan["C"], plan["C_model"], plan["heads"], dev
        self.softmax_scale = plan["softmax_scale"]
        with torch.cuda.device(dev):
            self.blocks = [tuple(_Layer(sp, dev, n_tile, cta_pair) for sp in blk) for blk in plan["blocks"]]
            self.convts = []
            for item in plan["convts"]:
                if item[0] == "conv":
                    self.convts.append(("conv", _Layer(item[1], dev, n_tile, cta_pair), item[2]))
                else:
                    self.convts.append(("col2im", _Layer(item[1], dev, n_tile, cta_pair), item[2], item[3].to(dev), item[4]))

    @classmethod
    def from_autoencoder(cls, autoencoder, precision: str = "fp32") -> "NativeDecoder":
        return cls(autoencoder.decoder, precision)

    max_elements = 2 ** 32 - 1   # the kernels form 32-bit element offsets into a layer's input

    def videos_per_pass(self, grid) -> int:
        """How many videos one pass can take: the largest layer input (the last transposed convolution's) must stay below
        `max_elements` elements; `forward_rows` splits larger batches (which also bounds the activation memory)."""
        rows = int(grid[0]) * int(grid[1]) * int(grid[2])
        worst = rows * 3 * self.C   # the attention output feeding `fc`
        for item in self.convts:
            worst = max(worst, rows * item[1].cin)
            rows *= item[2][0] * item[2][1] * item[2][2]
        return max(1, self.max_elements // worst)

    def forward_rows(self, x: torch.Tensor, B: int, grid) -> torch.Tensor:
        """Channels-last rows `[B * T * H * W, C]` -> video `[B, Cout, T', H', W']`."""
        T, H, W = (int(v) for v in grid)
        if max(T, H, W) > 32:
            raise D3PMError(f"latent grid {T}x{H}x{W}: the axial attention kernel covers axes of up to 32 positions")
        if x.shape[1] == self.C_model and self.C_model != self.C:   # zero channels up to the width the kernels work on
            x = torch.nn.functional.pad(x, (0, self.C - self.C_model))
        per_pass = self.videos_per_pass((T, H, W))
        if B > per_pass:
            n = T * H * W
            return torch.cat([self.forward_rows(x[b * n:min(B, b + per_pass) * n], min(B, b + per_pass) - b, (T, H, W))
                              for b in range(0, B, per_pass)], 0)
        lib = _lib.load_library()
        dev = x.device
        M = B * T * H * W
        for L3, L1, Lq, Lf in self.blocks:
            y = L1(L3(x, B, (T, H, W), terms=self.terms), B, (T, H, W), terms=self.terms)
            qkv = Lq(y, B, (T, H, W), terms=self.terms)
            att = torch.empty(M, 3 * self.C, dtype=torch.float32, device=dev)
            _lib.check(lib.d3pm_dec_axial_attention(qkv.data_ptr(), att.data_ptr(), B, T, H, W, self.heads, self.C // self.heads,
                                                    self.softmax_scale, ops._stream(dev)), "d3pm_dec_axial_attention")
            x = Lf(att, B, (T, H, W), terms=self.terms, residual=x)
        grid = (T, H, W)
        for item in self.convts:
            if item[0] == "conv":
                _, layer, stride = item
                x = layer(x, B, grid, terms=self.terms)
                grid = tuple(g * s for g, s in zip(grid, stride))
            else:
                _, layer, stride, bias, cout = item
                y_t = layer(x, B, grid, terms=self.terms, transposed=True)   # [B * T, 64 * cout, H * W]
                out = torch.empty(B, cout, *(g * s for g, s in zip(grid, stride)), dtype=torch.float32, device=dev)
                _lib.check(lib.d3pm_dec_col2im(y_t.data_ptr(), bias.data_ptr(), out.data_ptr(), B, *grid, cout, *stride,
                                               ops._stream(dev)), "d3pm_dec_col2im")
                return out
        raise D3PMError("decoder without transposed convolutions")

    def __call__(self, h: torch.Tensor) -> torch.Tensor:
        """Drop-in for `Decoder.forward(h)`: `h` is the reference's channels-first `[B, C, T, H, W]` tensor."""
        ops._need_cuda(h)
        if h.dim() != 5 or h.shape[1] != self.C_model:
            raise D3PMError(f"expected [B, {self.C_model}, T, H, W], got {tuple(h.shape)}")
        B, _, T, H, W = h.shape
        rows = h.detach().float().permute(0, 2, 3, 4, 1).reshape(B * T * H * W, self.C_model).contiguous()
        return self.forward_rows(rows, B, (T, H, W))


def decode(autoencoder, tokens: torch.Tensor, table: Optional[DecodeTable] = None, decoder: Optional[NativeDecoder] = None,
           status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Drop-in for `VQVAE.decode(encodings)` (videogpt_vq_vae.py:53-56).  With `decoder` (a `NativeDecoder` built from
    `autoencoder.decoder`) the whole chain runs on the library's kernels; without it the fused gather + 1x1x1 conv feeds the
    reference's own `decoder` module."""
    table = table if table is not None else DecodeTable.from_autoencoder(autoencoder)
    if decoder is None:
        return autoencoder.decoder(tokens_to_features(table, tokens, status))
    if tokens.dim() != 4:
        raise D3PMError("tokens must be [B, T', H', W'] (reshape the sampler's [B, N] with the latent grid, discrete_diffusion.py:62)")
    return decoder.forward_rows(embed_rows(table, tokens, status), tokens.shape[0], tokens.shape[1:])


def sample_and_decode(diffusion_model, autoencoder, text, text_emb: torch.Tensor, cf_text_emb: torch.Tensor, latent_shape,
                      table: Optional[DecodeTable] = None, decoder: Optional[NativeDecoder] = None) -> torch.Tensor:
    """The inference branch of the reference's caller, `DiscreteDiffusion.forward` (networks/discrete_diffusion.py:53-62):
    `sample(text, None, text_emb, cf_text_emb, content_token=None, filter_ratio=0)['content_token']`, viewed on the latent
    grid `(T', H', W')` (`.view(quant.shape)`, :62), then `autoencoder.decode`.  `diffusion_model` is the drop-in
    `FusedDiffusionTransformer` (or the reference's class: same `sample` signature)."""
    out = diffusion_model.sample(text, None, text_emb, cf_text_emb, content_token=None, filter_ratio=0)
    tokens = out["content_token"]
    return decode(autoencoder, tokens.view(tokens.shape[0], *latent_shape), table, decoder)

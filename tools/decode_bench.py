"""Token -> video at the shipped shape: the reference's `VQVAE.decode` in PyTorch on this GPU next to the native
chain (`decode.decode(..., decoder=NativeDecoder)`), same weights, CUDA events.

    python tools/decode_bench.py [--videos 16] [--layers]

n_hiddens 256, 3 residual blocks, downsample [1, 8, 8], 4 x 128 x 128 video, 4096 codes (ucf-ddiff-train.job:15).
`--layers` adds a per-launch table of the native chain (time, useful GFLOP, TFLOP/s).  Used by bench.py (`next_rows.decode`).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _timed(fn, n, dev, warm=2):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / n


def decoder_flops(B, grid, C, n_res, strides):
    """Useful multiply-adds x 2 of Decoder.forward (convolutions, Linears, attention products)."""
    T, H, W = grid
    M = B * T * H * W
    per_block = 2 * M * (27 * C * (C // 2) + (C // 2) * C + 9 * C * C + 3 * C * C) + 4 * M * (T + H + W) * C
    total = n_res * per_block
    g = [T, H, W]
    for i, s in enumerate(strides):
        cout = 3 if i == len(strides) - 1 else C
        taps = 1
        for d in s:
            taps *= 4 // d
        g = [a * b for a, b in zip(g, s)]
        total += 2 * B * g[0] * g[1] * g[2] * taps * C * cout
    return total


def run(videos=16, dev="cuda:0", reps=5, layers=False, quiet=False, n_tile=None, cta_pair=None):
    import torch
    from baseline import reference_loader as RL
    from d3pm_b200 import decode

    if not RL.reference_available():
        return {"skipped": "reference not staged on this box (baseline/_ref absent)"}
    torch.manual_seed(21)
    C, R, ds = 256, 3, (1, 8, 8)
    vq = RL.load_vqvae_module().VQVAE(checkpoint_path=None, embedding_dim=128, n_codes=4096, n_hiddens=C, n_res_layers=R,
                                      downsample=list(ds), sequence_length=4, resolution=128)
    with torch.no_grad():
        for m in vq.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    vq = vq.to(dev).eval()
    grid = (4, 16, 16)
    tokens = torch.randint(0, 4096, (videos, *grid), device=dev)
    table = decode.DecodeTable.from_autoencoder(vq)
    strides = [tuple(int(c.convt.stride[i]) for i in range(3)) for c in vq.decoder.convts]
    flops = decoder_flops(videos, grid, C, R, strides)
    res = {"what": f"VQVAE.decode of {videos} videos: tokens [{videos}, 4, 16, 16] -> video [{videos}, 3, 4, 128, 128]; n_hiddens 256, 3 attention "
                   f"residual blocks, 3 transposed convolutions (stride 1,2,2), fp32 weights, one B200", "useful_GFLOP": flops / 1e9}

    def ref_decode():
        with torch.no_grad():
            return vq.decode(tokens)

    saved = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        want = ref_decode()
        res["reference_fp32_ms"] = _timed(ref_decode, reps, dev)
        torch.backends.cudnn.allow_tf32 = True   # torch's defaults: cuDNN convolutions in TF32, matmuls in fp32
        res["reference_default_tf32_conv_ms"] = _timed(ref_decode, reps, dev)
        err_ref_tf32 = float((ref_decode() - want).abs().max())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    scale = float(want.abs().max())
    for prec in ("fp32", "tf32"):
        nd = decode.NativeDecoder(vq.decoder, precision=prec, n_tile=n_tile, cta_pair=cta_pair)
        got = decode.decode(vq, tokens, table, nd)
        ms = _timed(lambda: decode.decode(vq, tokens, table, nd), reps, dev)
        res[f"native_{prec}_ms"] = ms
        res[f"native_{prec}_useful_TFLOPs"] = flops / (ms * 1e-3) / 1e12
        res[f"native_{prec}_max_err_over_scale"] = float((got - want).abs().max()) / scale
        if layers:
            res[f"layers_{prec}"] = layer_table(nd, vq, tokens, table, dev)
    res["reference_default_tf32_conv_max_err_over_scale"] = err_ref_tf32 / scale
    res["speedup_fp32"] = res["reference_fp32_ms"] / res["native_fp32_ms"]
    res["speedup_tf32"] = res["reference_default_tf32_conv_ms"] / res["native_tf32_ms"]
    if not quiet:
        print(json.dumps(res, indent=1))
    return res


def layer_table(nd, vq, tokens, table, dev):
    """Per-launch times of one native decode (events around every d3pm_dec_conv call)."""
    import torch
    from d3pm_b200 import decode
    rows = []
    real = decode._Layer.__call__

    def timed_call(self, x, B, grid, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = real(self, x, B, grid, **kw)
        e1.record()
        rows.append((self, x.shape[0], e0, e1))
        return out

    decode._Layer.__call__ = timed_call
    try:
        decode.decode(vq, tokens, table, nd)
        rows.clear()
        decode.decode(vq, tokens, table, nd)
        torch.cuda.synchronize(dev)
    finally:
        decode._Layer.__call__ = real
    out = []
    for layer, m, e0, e1 in rows:
        ms = e0.elapsed_time(e1)
        fl = 2.0 * m * layer.nclass * layer.ntaps * layer.cin * layer.nout
        out.append({"M": m, "Cin": layer.cin, "taps": layer.ntaps, "classes": layer.nclass, "Nout": layer.nout, "n_tile": layer.n_tile,
                    "ms": ms, "GFLOP": fl / 1e9, "TFLOPs": fl / (ms * 1e-3) / 1e12})
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--layers", action="store_true")
    ap.add_argument("--n-tile", type=int, default=None, help="force the GEMM tile width (128 / 256) of every layer")
    ap.add_argument("--pairs", choices=["auto", "on", "off"], default="auto",
                    help="pairs of CTAs (cta_group::2) for the GEMMs: per layer (default), everywhere, nowhere")
    a = ap.parse_args()
    run(videos=a.videos, reps=a.reps, layers=a.layers, n_tile=a.n_tile, cta_pair={"auto": None, "on": True, "off": False}[a.pairs])

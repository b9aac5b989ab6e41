"""Token -> video, first stage: the fused codebook gather + 1x1x1 convolution against the reference's op
sequence `post_vq_conv(shift_dim(F.embedding(tokens, codebook), -1, 1))` (videogpt_vq_vae.py:53-56).  Needs a B200."""
import types

import pytest
import torch
import torch.nn.functional as F

from d3pm_b200 import decode, ops
from d3pm_b200._lib import D3PMError

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _autoencoder(K, E, C, seed):
    g = torch.Generator().manual_seed(seed)
    conv = torch.nn.Conv3d(E, C, 1)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(C, E, 1, 1, 1, generator=g) * 0.1)
        conv.bias.copy_(torch.randn(C, generator=g))
    ae = types.SimpleNamespace()
    ae.codebook = types.SimpleNamespace(embeddings=torch.randn(K, E, generator=g))
    ae.post_vq_conv = types.SimpleNamespace(conv=conv)
    ae.decoder = torch.nn.Identity()
    return ae


def _reference(ae, tokens):  # the reference's own op sequence, on the CPU
    h = F.embedding(tokens, ae.codebook.embeddings)
    h = h.permute(0, 4, 1, 2, 3).contiguous()  # shift_dim(h, -1, 1)
    return ae.post_vq_conv.conv(h)


@pytest.mark.parametrize("K,E,C,B,grid", [(4096, 128, 256, 2, (4, 16, 16)), (4096, 128, 256, 1, (16, 16, 16)), (512, 64, 240, 3, (3, 5, 7))])
def test_fused_gather_conv_matches_reference_ops(K, E, C, B, grid):
    ae = _autoencoder(K, E, C, 1)
    tokens = torch.randint(0, K, (B, *grid), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = _reference(ae, tokens)
    ae_dev = types.SimpleNamespace(codebook=types.SimpleNamespace(embeddings=ae.codebook.embeddings.to(DEV)),
                                   post_vq_conv=types.SimpleNamespace(conv=ae.post_vq_conv.conv.to(DEV)), decoder=torch.nn.Identity())
    status = ops.new_status(DEV)
    table = decode.DecodeTable.from_autoencoder(ae_dev)
    got = decode.tokens_to_features(table, tokens.to(DEV), status)
    assert got.shape == want.shape == (B, C, *grid)
    assert (got.cpu() - want).abs().max() <= 1e-5 * want.abs().max()
    assert int(status.item()) == 0
    assert torch.equal(decode.decode(ae_dev, tokens.to(DEV), table), got)  # decoder = Identity here


def test_mask_token_is_flagged():
    ae = _autoencoder(64, 16, 32, 3)
    ae.codebook.embeddings = ae.codebook.embeddings.to(DEV)
    ae.post_vq_conv.conv.to(DEV)
    table = decode.DecodeTable.from_autoencoder(ae)
    tokens = torch.full((1, 2, 2, 2), 64, dtype=torch.int64, device=DEV)  # [MASK] left in the grid
    status = ops.new_status(DEV)
    out = decode.tokens_to_features(table, tokens, status)
    assert int(status.item()) & 2 and float(out.abs().max()) == 0.0
    with pytest.raises(D3PMError):
        decode.DecodeTable(torch.zeros(64, 16, device=DEV), torch.zeros(32, 16, 3, 3, 3, device=DEV), None)


# ---- pinned to the reference itself: fixtures written by tests/golden/make_golden_decode.py from the imported VQVAE
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["decode_small", "decode_k512"])
def test_first_stage_against_reference_fixture(name):
    """`d3pm_decode_lut` + `d3pm_tokens_to_features` against `h`, the tensor the reference's own modules hand to its
    `Decoder` inside `VQVAE.decode` (videogpt_vq_vae.py:53-56)."""
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    table = decode.DecodeTable(torch.from_numpy(fx["codebook"]).to(DEV), torch.from_numpy(fx["conv_weight"]).to(DEV),
                               torch.from_numpy(fx["conv_bias"]).to(DEV))
    status = ops.new_status(DEV)
    got = decode.tokens_to_features(table, torch.from_numpy(fx["tokens"]).to(DEV), status).cpu()
    want = torch.from_numpy(fx["h"])
    assert got.shape == want.shape
    assert (got - want).abs().max() <= 2e-6 * max(1.0, float(want.abs().max()))
    assert int(status.item()) == 0


@pytest.mark.parametrize("name", ["decode_small", "decode_k512"])
def test_decode_with_the_reference_decoder_against_fixture(name):
    """`decode.decode(vq, tokens)` = fused first stage + the reference's own `Decoder` (PyTorch, on the GPU) with the
    fixture's weights, against the video `VQVAE.decode` returned on the CPU when the fixture was made."""
    from baseline import reference_loader as RL
    if not RL.reference_available():
        pytest.skip("reference not staged (baseline/_ref absent)")
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    E, K, H, R, d0, d1, d2, L, res, B = (int(v) for v in fx["hparams"])
    vq = RL.load_vqvae_module().VQVAE(checkpoint_path=None, embedding_dim=E, n_codes=K, n_hiddens=H, n_res_layers=R,
                                      downsample=[d0, d1, d2], sequence_length=L, resolution=res)
    vq.load_state_dict({k[3:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("sd/")}, strict=True)
    vq = vq.to(DEV).eval()
    tokens = torch.from_numpy(fx["tokens"]).to(DEV)
    # fp32 convolutions on both arms: torch lets cuDNN use TF32 by default, which would make the REFERENCE's first stage
    # (a 1x1x1 Conv3d) the less accurate of the two
    tf32_conv, tf32_mm = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            got = decode.decode(vq, tokens).cpu()
            ref_gpu = vq.decode(tokens).cpu()  # the reference's own decode on the same device
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_conv, tf32_mm
    want = torch.from_numpy(fx["video"])
    scale = float(want.abs().max())
    assert (got - ref_gpu).abs().max() <= 2e-5 * scale   # same decoder, same device: only the first stage differs
    assert (got - want).abs().max() <= 2e-4 * scale      # CPU fixture vs GPU convolutions

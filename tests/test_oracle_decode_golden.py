"""The decode fixtures (tests/golden/make_golden_decode.py) are what they claim: `h` is the embedding gather followed by the
1x1x1 convolution of the stored weights, and - where the reference tree is present - the live reference reproduces them."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from baseline import reference_loader as RL

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


ALL = ["decode_small", "decode_k512", "decode_h64", "decode_h128"]


@pytest.mark.parametrize("name", ALL)
def test_fixture_first_stage_is_gather_then_pointwise_conv(name):
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    emb = F.embedding(torch.from_numpy(fx["tokens"]), torch.from_numpy(fx["codebook"]))  # [B, T, H, W, E]
    w = torch.from_numpy(fx["conv_weight"]).flatten(1)                                    # [C, E]
    h = torch.einsum("bthwe,ce->bcthw", emb, w) + torch.from_numpy(fx["conv_bias"]).view(1, -1, 1, 1, 1)
    assert (h - torch.from_numpy(fx["h"])).abs().max() <= 1e-5


@pytest.mark.skipif(not RL.reference_available(), reason="reference tree not present on this machine")
@pytest.mark.parametrize("name", ALL)
def test_live_reference_reproduces_the_fixture(name):
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    E, K, H, R, d0, d1, d2, L, res, B = (int(v) for v in fx["hparams"])
    vq = RL.load_vqvae_module().VQVAE(checkpoint_path=None, embedding_dim=E, n_codes=K, n_hiddens=H, n_res_layers=R,
                                      downsample=[d0, d1, d2], sequence_length=L, resolution=res)
    # (the decode_h* fixtures keep the decode side of the state_dict only)
    vq.load_state_dict({k[3:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("sd/")}, strict=not name.startswith("decode_h"))
    with torch.no_grad():
        video = vq.eval().decode(torch.from_numpy(fx["tokens"]))
    assert torch.equal(video, torch.from_numpy(fx["video"]))


@pytest.mark.parametrize("name", ALL)
def test_decoder_oracle_reproduces_the_reference_video(name):
    """oracle/decoder_oracle.py (the restatement the CUDA decoder is checked against) vs the video the imported reference's
    `VQVAE.decode` returned when the fixture was made: same ops in the same order, fp32 on the CPU."""
    from oracle import decoder_oracle as DO
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    E, K, H, R, d0, d1, d2, L, res, B = (int(v) for v in fx["hparams"])
    sd = {k[3:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("sd/")}
    video = DO.vqvae_decode(sd, torch.from_numpy(fx["tokens"]), R, (d0, d1, d2))
    want = torch.from_numpy(fx["video"])
    assert video.shape == want.shape
    assert (video - want).abs().max() <= 2e-6 * max(1.0, float(want.abs().max()))
    dec = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    again = DO.decoder_forward(dec, torch.from_numpy(fx["h"]), R, DO.upsample_strides((d0, d1, d2)))
    assert (again - want).abs().max() <= 2e-6 * max(1.0, float(want.abs().max()))

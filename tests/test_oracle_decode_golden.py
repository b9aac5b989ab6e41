"""The decode fixtures (tests/golden/make_golden_decode.py) are what they claim: `h` is the embedding gather followed by the
1x1x1 convolution of the stored weights, and - where the reference tree is present - the live reference reproduces them."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from baseline import reference_loader as RL

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["decode_small", "decode_k512"])
def test_fixture_first_stage_is_gather_then_pointwise_conv(name):
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    emb = F.embedding(torch.from_numpy(fx["tokens"]), torch.from_numpy(fx["codebook"]))  # [B, T, H, W, E]
    w = torch.from_numpy(fx["conv_weight"]).flatten(1)                                    # [C, E]
    h = torch.einsum("bthwe,ce->bcthw", emb, w) + torch.from_numpy(fx["conv_bias"]).view(1, -1, 1, 1, 1)
    assert (h - torch.from_numpy(fx["h"])).abs().max() <= 1e-5


@pytest.mark.skipif(not RL.reference_available(), reason="reference tree not present on this machine")
@pytest.mark.parametrize("name", ["decode_small", "decode_k512"])
def test_live_reference_reproduces_the_fixture(name):
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    E, K, H, R, d0, d1, d2, L, res, B = (int(v) for v in fx["hparams"])
    vq = RL.load_vqvae_module().VQVAE(checkpoint_path=None, embedding_dim=E, n_codes=K, n_hiddens=H, n_res_layers=R,
                                      downsample=[d0, d1, d2], sequence_length=L, resolution=res)
    vq.load_state_dict({k[3:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("sd/")}, strict=True)
    with torch.no_grad():
        video = vq.eval().decode(torch.from_numpy(fx["tokens"]))
    assert torch.equal(video, torch.from_numpy(fx["video"]))

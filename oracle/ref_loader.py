"""Import the reference's own `diffusion_transformer.py` by path.

TEST INFRASTRUCTURE.  `/root/reference` exists only in the build container; `build()` stages the
few files of the path verbatim under the git-ignored `baseline/_ref/`, which travels to the GPU box
(`baseline/reference_loader.py` finds whichever is present).  Used by `tests/golden/make_golden*.py`
(to generate the committed fixtures) and by the cross-checks in `tests/`, which skip when neither
copy of the reference is present.

The reference module imports `hydra.utils.instantiate` (diffusion_transformer.py:16)
without using it; hydra is not installed here, so a two-attribute stub module is
registered before the import.  Bytecode writing is disabled so the read-only
reference tree stays pristine.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch

from baseline import reference_loader as _RL  # noqa: E402  (locates /root/reference or the staged baseline/_ref copy)


def reference_available() -> bool:
    return _RL.reference_available()


def load_reference_module():
    return _RL.load_diffusion_module()


class StubDenoiser(torch.nn.Module):
    """Stands in for `Text2ImageTransformer`: returns pre-generated logits.

    Exposes the three attributes `DiffusionTransformer` touches (`content_emb.num_embed`,
    `to_logits[-1].weight`, `forward(x_t, cond, t)`), and hands back the logits the way
    the real denoiser does: a `[B,K,N]` permuted view of a `[B,N,K]` tensor
    (transformer_utils.py:442-443).  Which tensor is returned is keyed on the first
    element of `cond` (1.0 -> conditional, 0.0 -> unconditional).
    """

    def __init__(self, num_codes: int, logits_c: torch.Tensor, logits_u: torch.Tensor):
        super().__init__()
        self.content_emb = types.SimpleNamespace(num_embed=num_codes + 1)
        self.to_logits = torch.nn.Sequential(torch.nn.Identity(), torch.nn.Linear(1, 1))
        self.logits_c, self.logits_u = logits_c, logits_u
        self.calls = 0

    def forward(self, x_t, cond, t):
        self.calls += 1
        src = self.logits_c if float(cond.flatten()[0]) > 0.5 else self.logits_u
        return src.permute(0, 2, 1)


def make_reference_model(num_codes: int, num_timesteps: int, seq_len: int, guidance_scale: float,
                         logits_c: torch.Tensor, logits_u: torch.Tensor):
    ref = load_reference_module()
    model = ref.DiffusionTransformer(
        transformer=StubDenoiser(num_codes, logits_c, logits_u),
        diffusion_step=num_timesteps, alpha_init_type="alpha1",
        guidance_scale=guidance_scale, content_seq_len=seq_len)
    return ref, model


@contextlib.contextmanager
def injected_uniform(u: torch.Tensor):
    """Make the next `torch.rand_like` calls return the shared uniform tensor `u`."""
    real = torch.rand_like

    def fake(x, *a, **k):
        assert x.shape == u.shape, (x.shape, u.shape)
        return u.to(x.dtype)

    torch.rand_like = fake
    try:
        yield
    finally:
        torch.rand_like = real

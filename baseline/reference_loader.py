"""Locate, stage and import the UNMODIFIED reference modules (never a copy that lives in git).

The reference (Developer-Zer0/GIF-synthesis-with-Discrete-Diffusion) is pure Python with dependencies that are not
installed anywhere here (hydra, pytorch_lightning, clip ...), so it cannot be pip-installed; but the few files on and
around the hot path import fine once two unused third-party names are stubbed:

    src/models/motionencoder/diffusion_transformer.py      the path itself (`DiffusionTransformer`)
    src/models/motionencoder/transformer_utils.py          the denoiser (`Text2ImageTransformer`), BASELINE config 3
    src/models/motionencoder/dalle_mask_image_embedding.py its token embedding
    src/models/networks/videogpt_vq_vae.py                 `VQVAE.decode`, the consumer of the tokens
    src/models/utils/model_utils.py                        `shift_dim` / `MultiHeadAttention` used by the VQ-VAE

`stage()` copies exactly these files, byte for byte, from `/root/reference` into `baseline/_ref/` (git-ignored, NOT
gpurun-ignored: it travels to the GPU box with the snapshot, like a built `.so`).  `reference_root()` searches
`$D3PM_REFERENCE_ROOT`, `/root/reference` (the build container) and `baseline/_ref` (the GPU box) in that order.

Who may use this: `bench.py --impl reference` (the reference arm), the config-3 block of `bench.py` / `tools/config3.py`
(reference `sample()` timed next to the drop-in), `tests/` and the golden-fixture generators.  The product package
(`d3pm_b200`) never imports it.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(HERE, "_ref")
SOURCE_ROOT = "/root/reference"

FILES = (
    "src/models/motionencoder/diffusion_transformer.py",
    "src/models/motionencoder/transformer_utils.py",
    "src/models/motionencoder/dalle_mask_image_embedding.py",
    "src/models/networks/videogpt_vq_vae.py",
    "src/models/utils/model_utils.py",
)


def stage(source_root: str = SOURCE_ROOT, dest_root: str = STAGED_ROOT) -> bool:
    """Copy the files above verbatim into `baseline/_ref/` (no-op when the source tree is absent)."""
    if not os.path.isfile(os.path.join(source_root, FILES[0])):
        return False
    for rel in FILES:
        dst = os.path.join(dest_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(source_root, rel), dst)
    return True


def reference_root():
    """First root that holds the reference's `diffusion_transformer.py`, or None."""
    for root in (os.environ.get("D3PM_REFERENCE_ROOT"), SOURCE_ROOT, STAGED_ROOT):
        if root and os.path.isfile(os.path.join(root, FILES[0])):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


def _stub_third_party():
    """`hydra.utils.instantiate` is imported and never called (diffusion_transformer.py:16, transformer_utils.py:15);
    `pytorch_lightning.LightningModule` is only a base class of VQVAE (videogpt_vq_vae.py:14)."""
    import torch

    if "hydra" not in sys.modules:
        hydra = types.ModuleType("hydra")
        hydra_utils = types.ModuleType("hydra.utils")
        hydra_utils.instantiate = lambda *a, **k: None
        hydra.utils = hydra_utils
        sys.modules["hydra"], sys.modules["hydra.utils"] = hydra, hydra_utils
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(torch.nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            @property
            def device(self):
                return next(self.parameters()).device

        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl


def _load(rel: str, name: str):
    root = reference_root()
    if root is None:
        raise FileNotFoundError(f"reference not found (looked in $D3PM_REFERENCE_ROOT, {SOURCE_ROOT}, {STAGED_ROOT}); "
                                f"run `python __graft_entry__.py build` where {SOURCE_ROOT} exists to stage it")
    sys.dont_write_bytecode = True  # keep the reference tree pristine
    _stub_third_party()
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_diffusion_module():
    """The reference's `diffusion_transformer` module (class `DiffusionTransformer`)."""
    return _load(FILES[0], "_d3pm_reference_diffusion_transformer")


def load_denoiser_modules():
    """(`transformer_utils`, `dalle_mask_image_embedding`) of the reference."""
    return _load(FILES[1], "_d3pm_reference_transformer_utils"), _load(FILES[2], "_d3pm_reference_dalle_embedding")


def load_vqvae_module():
    """The reference's `videogpt_vq_vae` module; it imports `src.models.utils.model_utils` by package path."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("reference not found")
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.append(root)  # at the END: the reference tree has its own `tests/` and `tools/` packages, which must not shadow ours
    return _load(FILES[3], "_d3pm_reference_videogpt_vq_vae")


def build_denoiser(num_codes: int, content_seq_len: int, spatial_size, *, n_layer=19, n_embd=64, n_head=16,
                   diffusion_step=100, condition_dim=512):
    """`Text2ImageTransformer` exactly as `configs/model/motionencoder/transformer_utils.yaml` +
    `dalle_mask_image_embedding.yaml` configure it (n_layer 19, n_embd 64, n_head 16, GELU2, adalayernorm, selfcross,
    mlp_hidden_times 4), sized for `content_seq_len` tokens (`spatial_size` H x W >= content_seq_len,
    dalle_mask_image_embedding.py:76-77)."""
    tu, de = load_denoiser_modules()
    dalle = de.DalleMaskImageEmbedding(num_embed=num_codes, spatial_size=list(spatial_size), embed_dim=n_embd,
                                       trainable=True, pos_emb_type="embedding")
    return tu.Text2ImageTransformer(
        dalle=dalle, attn_type="selfcross", n_layer=n_layer, condition_seq_len=77, content_seq_len=content_seq_len,
        content_spatial_size=list(spatial_size), n_embd=n_embd, condition_dim=condition_dim, n_head=n_head,
        attn_pdrop=0.0, resid_pdrop=0.0, block_activate="GELU2", timestep_type="adalayernorm", mlp_hidden_times=4,
        diffusion_step=diffusion_step)


def build_reference_model(transformer, *, diffusion_step=100, guidance_scale=2.0, content_seq_len=1024, **kw):
    """The reference's `DiffusionTransformer` around `transformer` (constructor of diffusion_transformer.py:72-164)."""
    ref = load_diffusion_module()
    return ref.DiffusionTransformer(transformer=transformer, diffusion_step=diffusion_step, alpha_init_type="alpha1",
                                    guidance_scale=guidance_scale, content_seq_len=content_seq_len, **kw)


@contextlib.contextmanager
def injected_uniform(next_uniform):
    """Inside the block every `torch.rand_like(x)` returns `next_uniform(x)`: the shared-noise hook of the parity
    tests (the reference draws its Gumbel noise from `torch.rand_like`, diffusion_transformer.py:355)."""
    import torch

    real = torch.rand_like

    def fake(x, *a, **k):
        return next_uniform(x).to(x.dtype)

    torch.rand_like = fake
    try:
        yield
    finally:
        torch.rand_like = real


if __name__ == "__main__":
    print("staged" if stage() else f"{SOURCE_ROOT} not present; nothing staged", "->", STAGED_ROOT)

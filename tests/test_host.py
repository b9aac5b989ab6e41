"""Host-side logic and the C-ABI surface, without a GPU."""
import ctypes
import os
import re
import subprocess
import types

import numpy as np
import pytest
import torch

import d3pm_b200
from d3pm_b200 import _lib, ops
from d3pm_b200.distributed import shard_range
from oracle import d3pm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_transformer(K):
    m = torch.nn.Module()
    m.content_emb = types.SimpleNamespace(num_embed=K + 1)
    return m


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "d3pm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(d3pm_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTED_SYMBOLS)
    lib = d3pm_b200.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", d3pm_b200.library_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (d3pm_[a-z0-9_]+)", nm))
    assert exported == set(declared)
    assert lib.d3pm_version() == 300


def test_argument_validation_happens_before_any_launch():
    lib = d3pm_b200.load_library()
    assert lib.d3pm_fused_step(None) == -1
    d = _lib.StepDesc()
    assert lib.d3pm_fused_step(ctypes.byref(d)) == -1
    assert b"required" in lib.d3pm_last_error()
    d.logits_c = d.x_t = d.t = d.coef_table = 16  # fake but non-null, aligned; rejected on the shape
    d.B, d.N, d.K, d.T = 1, 1, 6, 10
    assert lib.d3pm_fused_step(ctypes.byref(d)) == -3  # K % 4 != 0
    d.K, d.pitch_logits = 8, 9
    assert lib.d3pm_fused_step(ctypes.byref(d)) == -2  # pitch % 4 != 0
    d.pitch_logits, d.sample_mode = 8, 0
    assert lib.d3pm_fused_step(ctypes.byref(d)) == -1  # nothing to do
    assert lib.d3pm_build_coef_table(None, 10, 8, None, None) == -1
    assert lib.d3pm_q_posterior(None, 0, None, None, None, None, 0, 1, 1, 8, 10, None, None) == -1
    assert lib.d3pm_gumbel_argmax(None, 0, None, 0, 0, None, None, 1, 8, 0, 0, 0, None) == -1


def test_cpu_tensors_are_refused_loudly():
    with pytest.raises(d3pm_b200.D3PMError):
        ops.argmax_classes(torch.zeros(1, 4, 2))
    with pytest.raises(d3pm_b200.D3PMError):
        ops.build_coef_table(torch.zeros(8, 11), 10, 8)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setenv("D3PM_B200_LIB", "/nonexistent/libd3pm_b200.so")
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(d3pm_b200.D3PMError, match="no non-CUDA fallback"):
        _lib.load_library()


@pytest.mark.parametrize("T,K", [(100, 4096), (10, 64), (50, 2048), (200, 4096), (25, 64)])
def test_schedule_buffers_bit_identical_to_pinned_oracle(T, K):
    model = d3pm_b200.FusedDiffusionTransformer(transformer=_stub_transformer(K), diffusion_step=T,
                                                alpha_init_type="alpha1", guidance_scale=2, content_seq_len=8)
    sched = O.make_schedule(T, K)  # pinned bit-for-bit to the reference's buffers by make_golden.py
    for name in O.SCHEDULE_NAMES:
        assert torch.equal(getattr(model, name), sched[name]), name
    at, bt, ct, att, btt, ctt = d3pm_b200.alpha_schedule(T, N=K)
    assert at.shape == (T,) and att.shape == (T + 1,) and att[-1] == 1 and ctt[-1] == 0 and btt[-1] == 0


def test_module_surface_matches_reference():
    K = 64
    model = d3pm_b200.FusedDiffusionTransformer(transformer=_stub_transformer(K), diffusion_step=100,
                                                alpha_init_type="alpha1", guidance_scale=2, content_seq_len=16)
    keys = set(model.state_dict().keys())
    want = set(O.SCHEDULE_NAMES) | {"Lt_history", "Lt_count", "empty_text_embed"}
    assert want <= keys
    assert model.num_classes == K + 1 and model.num_timesteps == 100 and model.prior_rule == 0
    assert len(model.n_sample) == 100 and max(model.n_sample) <= 15
    for name in ("p_sample", "p_pred", "cf_predict_start", "predict_start", "q_posterior",
                 "log_sample_categorical", "sample", "update_n_sample"):
        assert callable(getattr(model, name))
    import inspect
    assert list(inspect.signature(model.p_sample).parameters) == ["log_x", "cond_emb", "cf_cond_emb", "t", "sampled", "to_sample"]
    assert list(inspect.signature(model.q_posterior).parameters) == ["log_x_start", "log_x_t", "t"]
    assert list(inspect.signature(model.sample).parameters)[:6] == [
        "condition_token", "condition_mask", "condition_embed", "cf_condition_embed", "content_token", "filter_ratio"]
    with pytest.raises(ValueError):
        d3pm_b200.FusedDiffusionTransformer(transformer=_stub_transformer(K), alpha_init_type="cos")


def test_reference_checkpoint_loads_strictly_both_ways():
    """The reference `DiffusionTransformer` (real `Text2ImageTransformer` inside) -> `state_dict()` -> the drop-in, strict,
    and back (diffusion_transformer.py:142-152 buffers + every denoiser parameter)."""
    from baseline import reference_loader as RL
    if not RL.reference_available():
        pytest.skip("reference not present (neither /root/reference nor baseline/_ref)")
    torch.manual_seed(0)
    ref = RL.build_reference_model(RL.build_denoiser(1024, 64, [8, 8]), content_seq_len=64)
    ours = d3pm_b200.FusedDiffusionTransformer(transformer=RL.build_denoiser(1024, 64, [8, 8]), diffusion_step=100,
                                                alpha_init_type="alpha1", guidance_scale=2.0, content_seq_len=64)
    sd = ref.state_dict()
    assert set(sd) == set(ours.state_dict())
    res = ours.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in ours.state_dict().items():
        assert torch.equal(v, sd[k]), k
    ref.load_state_dict(ours.state_dict(), strict=True)
    # the noise key is NOT part of the checkpoint contract; it has its own accessor
    state = ours.manual_seed(77, offset=5).rng_state()
    assert ours.manual_seed(1).set_rng_state(state).rng_state() == {"rng_seed": 77, "rng_offset": 5}


def test_layout_helpers_round_trip():
    rows = torch.arange(2 * 3 * 8, dtype=torch.float32).reshape(2, 3, 8)
    logical = ops.as_logical(rows, 5)
    assert logical.shape == (2, 5, 3) and logical[1, 4, 2] == rows[1, 2, 4]
    back, pitch = ops.rows_of(logical)
    assert pitch == 8 and back.data_ptr() == rows.data_ptr() and torch.equal(back, rows[:, :, :5])
    assert ops.rows_of(torch.zeros(2, 5, 3)) is None            # class-major (reference-contiguous) is not rows
    denoiser_out = torch.zeros(2, 3, 8).permute(0, 2, 1)        # what transformer_utils.py:442-443 returns
    assert ops.rows_of(denoiser_out)[1] == 8
    assert ops.padded_pitch(4097) == 4100 and ops.padded_pitch(65) == 68


def test_shard_range_partitions_the_batch():
    for B in (1, 7, 16, 128):
        for W in (1, 2, 3, 4, 8):
            spans = [shard_range(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)

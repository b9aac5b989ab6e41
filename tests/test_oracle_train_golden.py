"""Training-side oracle against the golden vectors generated from the reference (CPU)."""
import glob

import numpy as np
import pytest
import torch

from oracle import d3pm_oracle as O
from tests import helpers as H


def _lg(a):  # token-major numpy -> logical [B,C,N] tensor
    return torch.from_numpy(a).permute(0, 2, 1)


def test_forward_process_operators():
    fx = H.load(f"{H.GOLDEN}/forward_process_k64.npz")
    T, K = int(fx["T"]), int(fx["K"])
    sched = O.make_schedule(T, K)
    t = torch.from_numpy(fx["t"])
    soft, hot = _lg(fx["soft"]), O.index_to_log_onehot(torch.from_numpy(fx["x0"]), K + 1)
    hot_t = O.index_to_log_onehot(torch.from_numpy(fx["x_t"]), K + 1)
    assert torch.equal(O.q_pred(sched, soft, t), _lg(fx["qpred_soft"]))
    assert torch.equal(O.q_pred(sched, hot, t), _lg(fx["qpred_hot"]))
    assert torch.equal(O.q_pred(sched, hot, t - 1), _lg(fx["qpred_hot_tm1"]))   # t = 0 wraps to the identity slot
    assert torch.equal(O.q_pred_one_timestep(sched, hot_t, t), _lg(fx["qone_hot"]))
    assert torch.equal(O.q_pred_one_timestep(sched, soft, t), _lg(fx["qone_soft"]))
    xs = O.q_sample(sched, hot, t, _lg(fx["uniform"])).argmax(1)
    assert np.array_equal(xs.numpy(), fx["q_sample"])


@pytest.mark.parametrize("path", sorted(glob.glob(f"{H.GOLDEN}/train_loss_*.npz")), ids=lambda p: p.split("train_loss_")[-1][:-4])
def test_train_loss_forward_and_backward(path):
    fx = H.load(path)
    T, K = int(fx["T"]), int(fx["K"])
    sched = O.make_schedule(T, K)
    logits = torch.from_numpy(fx["logits"]).clone().requires_grad_(True)   # [B,N,K], the denoiser's physical layout
    x0, t, pt = torch.from_numpy(fx["x0"]), torch.from_numpy(fx["t"]), torch.from_numpy(fx["pt"])
    lmp, vb, x0r, xt, _ = O.train_loss(sched, logits.permute(0, 2, 1), x0, t, pt, _lg(fx["uniform"]),
                                       auxiliary_loss_weight=float(fx["aux"]), adaptive_auxiliary_loss=bool(fx["adaptive"]))
    assert np.array_equal(xt.numpy(), fx["x_t"]) and np.array_equal(x0r.numpy(), fx["x0_recon"])
    assert torch.equal(lmp.detach(), _lg(fx["log_model_prob"]))
    assert np.array_equal(vb.detach().numpy(), fx["vb_loss"])
    loss = vb.sum() / (x0.shape[0] * x0.shape[1])                          # forward() (:554)
    assert float(loss) == float(fx["loss"])
    loss.backward()
    assert np.abs(logits.grad.numpy() - fx["grad_logits"]).max() <= 1e-7 * max(1.0, np.abs(fx["grad_logits"]).max())

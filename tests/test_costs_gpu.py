"""Functional forms of the reference's loss helpers (multimodal_segmentation_b200/costs.py <- reference costs.py:24-143)
on CUDA tensors, against the oracle (oracle/ref_ops.py, itself pinned to the reference's own costs.py through
tests/golden/golden_ref.npz).  They are thin wrappers over the loss kernels the trainers use; fp32 bound 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from tests.util import rel_l2

pytestmark = [pytest.mark.gpu]


def _pair(seed, B=3, H=20, W=24, C=5):
    rs = np.random.RandomState(seed)
    logits = rs.normal(size=(B, H, W, C)).astype(np.float32)
    pred = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    lab = rs.randint(0, C, size=(B, H, W))
    true = np.eye(C, dtype=np.float32)[lab]
    return true, pred.astype(np.float32)


def _g(a):
    return torch.from_numpy(a).cuda()


def _t(a):
    return torch.from_numpy(a).double()


def test_dice_helpers():
    from multimodal_segmentation_b200 import costs
    true, pred = _pair(0)
    got = costs.dice_coef_perbatch(_g(true), _g(pred)).cpu().numpy()
    assert rel_l2(got, R.dice_coef_perbatch(_t(true), _t(pred)).numpy()) < 1e-4
    got = costs.dice_coef_loss(_g(true), _g(pred)).item()
    assert abs(got - R.dice_coef_perbatch(_t(true), _t(pred)).mean().item()) < 1e-4
    got = costs.make_dice_loss_fnc(4)(_g(true), _g(pred)).item()
    assert abs(got - R.dice_loss(_t(true), _t(pred), 4).item()) < 1e-4


def test_cross_entropy_helpers_keep_the_reference_argument_order():
    from multimodal_segmentation_b200 import costs
    true, pred = _pair(1)
    # signature (y_pred, y_true): counts from the second argument, log of the first
    ref = R.weighted_cross_entropy_loss(_t(pred), _t(true)).item()
    got = costs.weighted_cross_entropy_loss(_g(pred), _g(true)).item()
    assert abs(got - ref) < 1e-4 * max(1.0, abs(ref))
    ref = R.weighted_cross_entropy_perbatch(_t(pred), _t(true)).numpy()
    got = costs.weighted_cross_entropy_perbatch(_g(pred), _g(true)).cpu().numpy()
    assert rel_l2(got, ref) < 1e-4
    # the combined losses call them with swapped arguments (costs.py:134,142)
    ref = R.combined_dice_bce(_t(true), _t(pred), 4).item()
    got = costs.make_combined_dice_bce(4)(_g(true), _g(pred)).item()
    assert abs(got - ref) < 1e-4 * max(1.0, abs(ref))
    ref = R.combined_dice_bce_perbatch(_t(true), _t(pred), 4).numpy()
    got = costs.make_combined_dice_bce_perbatch(4)(_g(true), _g(pred)).cpu().numpy()
    assert rel_l2(got, ref) < 1e-4


def test_mae_single_input():
    from multimodal_segmentation_b200 import costs
    rs = np.random.RandomState(2)
    a, b = (rs.normal(size=(3, 20, 24, 1)).astype(np.float32) for _ in range(2))
    got = costs.mae_single_input([_g(a), _g(b)]).cpu().numpy()
    ref = R.mae_single_input(_t(a), _t(b)).numpy().reshape(-1, 1)
    assert got.shape == (3, 1) and rel_l2(got, ref) < 1e-4

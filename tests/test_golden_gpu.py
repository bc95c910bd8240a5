"""The CUDA kernels (through the C-ABI) against golden vectors produced by the reference's own Python source
(tests/golden/make_golden.py -> tests/golden/golden_ref.npz).  fp32 kernels vs fp64 goldens: tolerances are
fp32 accumulation noise; rounding masks are bit-exact."""
import os

import numpy as np
import pytest
import torch

from tests.util import cpu, gpu, rel_l2

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ref.npz"))


@pytest.fixture(scope="module")
def ops():
    from multimodal_segmentation_b200 import ops as o
    return o


def test_rounding_bit_exact(ops):
    y = ops.round_fwd(gpu(G["round_x"]))
    assert np.array_equal(cpu(y), G["round_y"])


def test_film_and_resize(ops):
    y = ops.film_fwd(gpu(G["film_x"]), gpu(G["film_gamma"]), gpu(G["film_beta"]))
    assert rel_l2(cpu(y), G["film_y"]) < 1e-6
    assert np.array_equal(cpu(ops.resize_nn_fwd(gpu(G["resize_in"]), 6, 4)), G["resize_6x4"].astype(np.float32))
    assert np.array_equal(cpu(ops.resize_nn_fwd(gpu(G["resize_in"]), 3, 2)), G["resize_3x2"].astype(np.float32))


def test_tps_fast_path_matches_reference_layer(ops):
    """fused spline-evaluate + bilinear kernel == ThinPlateSpline2D.call of the reference (inverse=False)"""
    out, locs = ops.tps_warp_fwd(gpu(G["tps_vol"]), gpu(G["spline_theta"]), (5, 5), want_locs=True)
    H, W = G["tps_vol"].shape[1:3]
    ref_locs = G["tps_locs_inv0"][..., ::-1] * np.array([W - 1, H - 1])      # (row,col) normalised -> (x,y) pixels
    assert np.abs(cpu(locs) - ref_locs).max() < 2e-4                           # pixels
    assert rel_l2(cpu(out), G["tps_warped_inv0"]) < 1e-4


def test_tps_general_path_matches_reference(ops):
    """batched LU solve + apply kernels == layers/interpolate_spline.py (order 2, the layer's usage and inverse=True)"""
    cp = np.repeat(G["spline_cp"], 3, 0)
    warped = cp + G["spline_theta"]
    q = np.repeat(G["spline_q"], 3, 0)
    w, v = ops.tps_solve(gpu(cp), gpu(warped))
    out = ops.tps_apply(gpu(q), gpu(cp), w, v)
    assert np.abs(cpu(out) - G["spline_order2"]).max() < 2e-4
    w, v = ops.tps_solve(gpu(warped), gpu(cp))          # inverse=True: train on the warped grid
    out = ops.tps_apply(gpu(q), gpu(warped), w, v)
    assert np.abs(cpu(out) - G["tps_locs_inv1"]).max() < 2e-3
    for key, order, reg in (("spline_order1", 1, 0.0), ("spline_order4_reg", 4, 0.01), ("spline_order2_reg", 2, 0.003)):
        w, v = ops.tps_solve(gpu(G["spline_tp"]), gpu(G["spline_tv"]), order, reg)
        out = ops.tps_apply(gpu(G["spline_qq"]), gpu(G["spline_tp"]), w, v, order)
        assert rel_l2(cpu(out), G[key]) < 5e-3, key


def test_losses_match_reference_costs(ops):
    pred, true = G["loss_pred"], G["loss_true"]
    loss = ops.zeros(1)
    ops.segloss(gpu(pred), gpu(true), 4, 1, 1.0, loss, want_grad=False)       # dice(4 ch) + 0.01 * swapped wBCE(5 ch)
    assert abs(float(cpu(loss)[0]) - float(G["combined_dice_bce4"])) < 1e-4 * max(1.0, abs(float(G["combined_dice_bce4"])))
    loss = ops.zeros(1)
    ops.segloss(gpu(pred), gpu(true), 4, 0, 1.0, loss, want_grad=False)
    assert abs(float(cpu(loss)[0]) - float(G["dice_fnc4"])) < 1e-5
    loss = ops.zeros(1)
    z, klv = ops.vae_fwd(gpu(G["kl_mu"]), gpu(G["kl_lv"]), gpu(G["sampling_eps"]), 1.0, loss)
    assert rel_l2(cpu(klv), G["kl"]) < 1e-5
    assert rel_l2(cpu(z), G["sampling_z"]) < 1e-5
    assert abs(float(cpu(loss)[0]) - float(G["kl"].mean())) < 1e-4


def test_spectral_regulariser_matches_reference(ops):
    W = G["spectral_W"]
    dim, cout = W.shape[0] * W.shape[1] * W.shape[2], W.shape[3]
    loss = ops.zeros(1)
    dW = ops.zeros(dim, cout)
    ops.spectral_reg(gpu(W).view(dim, cout), gpu(G["spectral_u0"]), 10.0, loss, dW)
    assert abs(float(cpu(loss)[0]) - float(G["spectral_loss"])) < 1e-4 * float(G["spectral_loss"])


def test_automated_pairing_kernels_against_reference_vectors(ops):
    """csrc/pairing.cu forward values against the vectors produced by the reference's costs.py / balancer.py"""
    true5, pred = gpu(G["loss_true"]), gpu(G["loss_pred"])
    B = true5.shape[0]
    L = torch.empty(2, B, device="cuda")
    ops.segloss_pb_fwd(pred, true5, 4, L[0])
    ops.mae_pb_fwd(gpu(G["pb_mae_y"]), gpu(G["pb_mae_x"]), L[1])
    assert rel_l2(cpu(L[0]), G["pb_combined5"]) < 1e-5
    assert rel_l2(cpu(L[1]), G["pb_mae"][:, 0]) < 1e-5
    d = ops.pair_dice(gpu(G["bal_a"]), gpu(G["bal_b"]))
    assert rel_l2(cpu(d), G["bal_dice"]) < 1e-6

"""The oracle against golden vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/make_golden.py runs
layers/interpolate_spline.py, layers/stn_spline.py, layers/film.py, layers/spade.py, layers/rounding.py,
layers/spectralnorm.py, costs.py, utils/*.py, model_executors/base_executor.py of /root/reference on numpy
stand-ins for the TF/Keras entry points they call).  The host-side helpers of the product (rescale, sample,
add_residual, align_batches, nDgrid, NormalDistribution) are checked against the same file."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ref.npz"))


def T(a):
    return torch.from_numpy(np.asarray(a, np.float64))


def close(a, b, tol=1e-10):
    a = np.asarray(a.detach() if hasattr(a, "detach") else a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), np.abs(a - b).max()


def test_ndgrid():
    close(R.nDgrid((5, 5), torch.float64), G["ndgrid_5x5"], 1e-7)          # oracle grid is the float32 cast
    from multimodal_segmentation_b200.layers.stn_spline import nDgrid
    close(nDgrid([5, 5], dtype="float64"), G["ndgrid_5x5"])
    close(nDgrid([3, 4], normalise=False, dtype="float64"), G["ndgrid_3x4_unnorm"])
    close(nDgrid([3, 4], center=True, dtype="float64"), G["ndgrid_3x4_center"])


def test_interpolate_spline_layer_usage():
    cp, q, theta = T(G["spline_cp"]), T(G["spline_q"]), T(G["spline_theta"])
    for b in range(3):
        out = R.interpolate_spline(cp, cp + theta[b:b + 1], q, 2)
        close(out, G["spline_order2"][b:b + 1], 1e-9)
    w, v = R.solve_interpolation(cp, cp + theta[0:1], 2)
    close(w, G["spline_order2_w0"], 1e-8)
    close(v, G["spline_order2_v0"], 1e-8)
    # float32 execution (what TF would run) stays within fp32 noise of the fp64 oracle
    out32 = torch.cat([R.interpolate_spline(cp.float(), (cp + theta[b:b + 1]).float(), q.float(), 2) for b in range(3)], 0)
    close(out32.double(), G["spline_order2_f32"], 2e-4)


@pytest.mark.parametrize("key,order,reg", [("spline_order1", 1, 0.0), ("spline_order4_reg", 4, 0.01),
                                            ("spline_order2_reg", 2, 0.003), ("spline_order3", 3, 0.0)])
def test_interpolate_spline_general(key, order, reg):
    out = R.interpolate_spline(T(G["spline_tp"]), T(G["spline_tv"]), T(G["spline_qq"]), order, reg)
    close(out, G[key], 1e-8)


@pytest.mark.parametrize("inv", [0, 1])
def test_thin_plate_spline_layer(inv):
    vol, theta = T(G["tps_vol"]), T(G["spline_theta"])
    out = R.thin_plate_spline_2d(vol, theta, (5, 5), 2, inverse=bool(inv))
    out = out[0] if isinstance(out, tuple) else out
    close(out, G["tps_warped_inv%d" % inv], 1e-6)      # oracle control grid is a float32 cast (as in the reference)


def test_film_spade_resize_round():
    close(R.film(T(G["film_x"]), T(G["film_gamma"]), T(G["film_beta"])), G["film_y"])
    close(R.spade_cond(T(G["film_x"]), T(G["spade_gamma"]), T(G["spade_beta"])), G["spade_y"])
    close(R.resize_nn(T(G["resize_in"]), 6, 4), G["resize_6x4"])
    close(R.resize_nn(T(G["resize_in"]), 3, 2), G["resize_3x2"])
    y = R.rounding(torch.from_numpy(G["round_x"]))
    assert np.array_equal(y.numpy(), G["round_y"])        # bit exact


def test_losses():
    pred, true = T(G["loss_pred"]), T(G["loss_true"])
    close(R.dice_loss(true, pred, 4), G["dice_fnc4"])
    close(R.dice_coef_perbatch(true[..., :4], pred[..., :4]), G["dice_perbatch4"])
    close(R.weighted_cross_entropy_loss(true, pred), G["wbce_as_called"], 1e-9)     # swapped arguments, as called
    close(R.combined_dice_bce(true, pred, 4), G["combined_dice_bce4"], 1e-9)
    close(R.kl(T(G["kl_mu"]), T(G["kl_lv"])), G["kl"])
    close(R.sampling(T(G["kl_mu"]), T(G["kl_lv"]), T(G["sampling_eps"])), G["sampling_z"])
    assert abs(R.np_dice(G["loss_true"][..., :4], G["loss_pred"]) - G["dice_metric"]) < 1e-12
    assert abs(R.np_dice(G["loss_true"][..., :4], G["loss_pred"], binarise=True) - G["dice_metric_bin"]) < 1e-12


def test_spectral_regulariser():
    loss = R.spectral_reg(T(G["spectral_W"]), T(G["spectral_u0"]), 10.0)
    close(loss, G["spectral_loss"], 1e-9)


def test_host_helpers():
    from multimodal_segmentation_b200.utils import data_utils
    from multimodal_segmentation_b200.utils.distributions import NormalDistribution
    from multimodal_segmentation_b200.model_executors.base_executor import Executor
    close(data_utils.rescale(G["rescale_in"].copy(), -1, 1), G["rescale_out"])
    close(data_utils.rescale(np.full((1, 4, 4, 1), 3.0), -1, 1), G["rescale_const"])
    assert np.array_equal(data_utils.sample(np.arange(20).reshape(10, 2), 4, seed=5), G["sample_seed5"])
    assert np.array_equal(Executor.add_residual(None, G["residual_in"]), G["residual_out"])
    al = Executor.align_batches(None, [np.arange(10.).reshape(5, 2), np.arange(6.).reshape(3, 2)])
    assert np.array_equal(al[0], G["align_0"]) and np.array_equal(al[1], G["align_1"])
    np.random.seed(3)
    close(NormalDistribution().sample((3, 8)), G["normal_dist_seed3"])


def test_automated_pairing_losses():
    """costs.make_combined_dice_bce_perbatch / weighted_cross_entropy_perbatch / mae_single_input and the Balancer's
    Dice overlap, with the argument order of models/dafnet.py:283-315"""
    true5, pred = T(G["loss_true"]), T(G["loss_pred"])
    close(R.combined_dice_bce_perbatch(true5, pred, 4), G["pb_combined5"])
    close(R.weighted_cross_entropy_perbatch(true5, pred), G["pb_wce_as_called"])
    close(R.mae_single_input(T(G["pb_mae_x"]), T(G["pb_mae_y"])), G["pb_mae"])
    close(R.pair_dice(T(G["bal_a"]), T(G["bal_b"])), G["bal_dice"])

"""Pins the CPU oracle (oracle/ref_ops.py).

The reference ships no tests, fixtures or golden vectors and its TF 1.4 stack cannot be
installed here, so the oracle is pinned by:
  * known-answer tests that follow from the reference code alone,
  * independent cross-oracles: scipy RBFInterpolator (thin-plate spline), torch grid_sample
    (tf.contrib.resampler), np.round, a second pure-numpy loop convolution,
  * fp64 finite-difference checks of the hand-derived gradient conventions.
CPU only; runs in seconds.
"""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R

torch.manual_seed(0)


def rng(s):
    return np.random.RandomState(s)


def t64(a, grad=False):
    r = torch.as_tensor(np.asarray(a, np.float64))
    if grad:
        r.requires_grad_(True)
    return r


# ------------------------------------------------------------------ KATs from the reference code
def test_rounding_half_to_even_and_ste():
    x = torch.tensor([0.5, 1.5, 2.5, -0.5, 0.49999, 0.50001], requires_grad=True)
    y = R.rounding(x)
    assert y.tolist() == [0.0, 2.0, 2.0, -0.0, 0.0, 1.0]
    y.sum().backward()
    assert x.grad.tolist() == [1.0] * 6          # layers/rounding.py:40-42


def test_rounded_softmax_is_at_most_one_hot():
    p = R.softmax(torch.randn(1000, 8) * 4)
    r = R.rounding(p)
    assert r.sum(-1).max() <= 1.0
    assert set(r.unique().tolist()) <= {0.0, 1.0}


def test_film_and_spade_identities():
    x = torch.randn(2, 5, 5, 8)
    assert torch.equal(R.film(x, torch.ones(2, 8), torch.zeros(2, 8)), x)            # layers/film.py:36
    assert torch.equal(R.spade_cond(x, torch.zeros_like(x), torch.zeros_like(x)), x)  # layers/spade.py:55


def test_dice_kl_kats():
    m = torch.zeros(2, 6, 6, 5)
    m[..., 0] = 1
    assert abs(R.dice_loss(m, m, 4).item()) < 1e-9                   # dice(x,x) -> loss 0 (costs.py:43-48)
    assert torch.all(R.kl(torch.zeros(3, 8), torch.zeros(3, 8)) == 0)  # costs.py:186-189
    assert R.np_dice(m.numpy(), m.numpy(), binarise=True) == pytest.approx(1.0)


def test_weighted_bce_argument_swap():
    """costs.py:134 passes (y_true, y_pred) into a (y_pred, y_true) signature: class counts come from
    the prediction and log() is taken of the target -> 27.631*mean(sum_c pred_c*(1-target_c)*w_c)."""
    r = rng(0)
    pred = R.softmax(torch.as_tensor(r.normal(size=(2, 4, 4, 5)))).double()
    tgt = torch.as_tensor(np.eye(5)[r.randint(0, 5, size=(2, 4, 4))])
    got = R.weighted_cross_entropy_loss(tgt, pred)        # as called by combined_dice_bce
    n = pred.sum(dim=(0, 1, 2))
    w = n.sum() / (n + 1e-12)
    expect = (-(pred * torch.log(tgt + 1e-12) * w).sum(-1)).mean()
    assert got.item() == pytest.approx(expect.item(), rel=1e-12)
    closed = -np.log(1e-12) * ((pred * (1 - tgt) * w).sum(-1)).mean()
    assert got.item() == pytest.approx(closed.item(), rel=1e-9)


def test_tps_identity_and_affine_reproduction():
    """theta = 0 -> identity warp (stn_spline.py:116); a spline reproduces train_values at
    train_points and any affine map exactly (interpolate_spline.py:225-227)."""
    vol = torch.rand(2, 12, 10, 3, dtype=torch.float64)
    out, locs = R.thin_plate_spline_2d(vol, torch.zeros(2, 25, 2, dtype=torch.float64))
    assert (out - vol).abs().max() < 1e-6
    c = R.nDgrid((5, 5), torch.float64)
    A = torch.tensor([[1.1, 0.2], [-0.3, 0.9]], dtype=torch.float64)
    f = c @ A + 0.05
    q = torch.rand(1, 50, 2, dtype=torch.float64)
    got = R.interpolate_spline(c, f, q, 2)
    assert (got - (q @ A + 0.05)).abs().max() < 1e-9
    assert (R.interpolate_spline(c, f, c, 2) - f).abs().max() < 1e-9


# ------------------------------------------------------------------ independent cross-oracles
def test_spline_matches_scipy_rbf():
    from scipy.interpolate import RBFInterpolator
    r = rng(1)
    c = R.nDgrid((5, 5), torch.float64)[0].numpy() + r.normal(size=(25, 2)) * 0.01
    f = r.normal(size=(25, 2))
    q = r.uniform(0, 1, size=(200, 2))
    ours = R.interpolate_spline(t64(c)[None], t64(f)[None], t64(q)[None], 2)[0].numpy()
    # scipy's kernel is r^2 log r = 2 * (0.5 r^2 log r^2 / 2): same interpolant (weights rescale)
    ref = RBFInterpolator(c, f, kernel="thin_plate_spline", degree=1)(q)
    assert np.abs(ours - ref).max() < 1e-9


def test_resampler_matches_grid_sample():
    r = rng(2)
    B, H, W, C, m = 2, 7, 9, 3, 400
    vol = t64(r.normal(size=(B, H, W, C)))
    warp = t64(np.stack([r.uniform(-2, W + 1, (B, m)), r.uniform(-2, H + 1, (B, m))], -1))
    ours = R.resampler(vol, warp)
    grid = torch.stack([2 * warp[..., 0] / (W - 1) - 1, 2 * warp[..., 1] / (H - 1) - 1], -1)[:, :, None, :]
    ref = torch.nn.functional.grid_sample(vol.permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="zeros",
                                          align_corners=True)[:, :, :, 0].permute(0, 2, 1)
    assert (ours - ref).abs().max() < 1e-12


@pytest.mark.parametrize("k,s,p", [(3, 1, 1), (3, 2, 0), (5, 1, 0), (4, 2, 0), (1, 1, 0)])
def test_conv_two_statements_agree(k, s, p):
    r = rng(k + s)
    x, w, b = r.normal(size=(2, 11, 9, 3)), r.normal(size=(k, k, 3, 4)), r.normal(size=4)
    a = R.conv2d(t64(x), t64(w), t64(b), s, "same" if p else "valid").numpy()
    assert np.abs(a - R.conv2d_loops(x, w, b, s, p)).max() < 1e-12


def test_batchnorm_pool_second_statement():
    r = rng(3)
    x = r.normal(size=(3, 6, 8, 5))
    y, mean, var = R.batchnorm_train(t64(x), torch.ones(5, dtype=torch.float64), torch.zeros(5, dtype=torch.float64))
    xm = x.reshape(-1, 5)
    assert np.abs(mean.numpy() - xm.mean(0)).max() < 1e-12 and np.abs(var.numpy() - xm.var(0)).max() < 1e-12
    assert np.abs(y.numpy() - (x - xm.mean(0)) / np.sqrt(xm.var(0) + 1e-3)).max() < 1e-12
    mp = R.maxpool2(t64(x)).numpy()
    loops = np.zeros((3, 3, 4, 5))
    for i in range(3):
        for j in range(4):
            loops[:, i, j] = x[:, 2 * i:2 * i + 2, 2 * j:2 * j + 2].max(axis=(1, 2))
    assert np.array_equal(mp, loops)
    up = R.upsample2(t64(x)).numpy()
    assert np.array_equal(up[:, ::2, ::2], x) and np.array_equal(up[:, 1::2, 1::2], x)
    # instance norm over H,W,C jointly with eps added to the std (keras_contrib, axis=None)
    inn = R.instance_norm_axis_none(t64(x)).numpy()
    for b in range(3):
        assert np.abs(inn[b] - (x[b] - x[b].mean()) / (x[b].std() + 1e-3)).max() < 1e-12


def test_resize_nearest_picks_top_left():
    x = torch.arange(2 * 8 * 8 * 1, dtype=torch.float64).reshape(2, 8, 8, 1)
    y = R.resize_nn(x, 4, 4)
    assert torch.equal(y, x[:, ::2, ::2])


# ------------------------------------------------------------------ gradient conventions
def test_leaky_relu_and_maximum_gradient_conventions():
    x = torch.tensor([-1.0, 0.0, 2.0], requires_grad=True)
    R.leaky_relu(x, 0.3).sum().backward()
    assert x.grad.tolist() == pytest.approx([0.3, 0.0, 1.0])        # derivative at exactly 0 is 0
    a = torch.tensor([1.0, 0.0, 0.0], requires_grad=True)
    b = torch.tensor([1.0, 0.0, 1.0], requires_grad=True)
    R.tf_maximum(a, b).sum().backward()
    assert a.grad.tolist() == [1.0, 1.0, 0.0] and b.grad.tolist() == [0.0, 0.0, 1.0]   # ties -> first input


def _fd(f, x, eps=1e-6):
    g = np.zeros_like(x)
    for i in np.ndindex(*x.shape):
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        g[i] = (f(xp) - f(xm)) / (2 * eps)
    return g


def test_finite_difference_tps_theta_gradient():
    r = rng(4)
    vol = t64(r.uniform(size=(1, 8, 8, 2)))
    theta = r.normal(size=(1, 25, 2)) * 0.02
    g = t64(r.normal(size=(1, 8, 8, 2)))
    th = t64(theta, grad=True)
    (R.thin_plate_spline_2d(vol, th)[0] * g).sum().backward()
    fd = _fd(lambda a: float((R.thin_plate_spline_2d(vol, t64(a))[0] * g).sum()), theta)
    assert np.abs(fd - th.grad.numpy()).max() < 1e-5


def test_finite_difference_losses():
    r = rng(5)
    pred = R.softmax(t64(r.normal(size=(2, 3, 3, 5)))).numpy()
    tgt = t64(np.eye(5)[r.randint(0, 5, size=(2, 3, 3))])
    p = t64(pred, grad=True)
    R.combined_dice_bce(tgt, p, 4).backward()
    fd = _fd(lambda a: float(R.combined_dice_bce(tgt, t64(a), 4)), pred)
    assert np.abs(fd - p.grad.numpy()).max() < 1e-5


def test_spectral_reg_value_and_gradient():
    r = rng(6)
    W = r.normal(size=(4, 4, 3, 5)) * 0.3
    u0 = r.uniform(-1, 1, size=(48, 1))
    w = t64(W, grad=True)
    loss = R.spectral_reg(w, t64(u0), 10.0)
    loss.backward()
    x = W.reshape(-1, 5)
    sigma = np.linalg.svd(x, compute_uv=False)[0]
    # three power iterations get close to the top singular value
    approx = 10.0 * np.abs(x / sigma - x).mean()
    assert loss.item() == pytest.approx(approx, rel=0.2)
    s = np.sign(x / sigma - x)
    assert np.mean(np.sign(-w.grad.numpy().reshape(-1, 5)) == s) > 0.95


def test_adam_matches_closed_form_first_step():
    p, g = np.array([1.0, -2.0]), np.array([0.5, -0.25])
    p1, m1, v1 = R.adam_step(p, g, np.zeros(2), np.zeros(2), 1)
    # first step of Adam moves every weight by ~lr*sign(g)
    assert np.allclose(p1, p - 1e-4 * np.sign(g), atol=1e-9)

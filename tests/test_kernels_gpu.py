"""Parity of every C-ABI kernel (through ctypes -> libdafk.so) against the CPU oracle.

Integer-like results (rounding masks, argmax routing) are compared bit-exactly; floating point
within the tolerance stated next to each assert (fp32 kernels: <= 1e-4 relative L2, the
north-star bound; most are ~1e-6).
"""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from tests.util import cpu, gpu, rel_l2, t

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    from multimodal_segmentation_b200 import ops as o
    return o


def rng(seed):
    return np.random.RandomState(seed)


# ------------------------------------------------------------------ rounding / softmax
@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 1 << 20, (1 << 20) + 3])
def test_round_bit_exact(ops, n):
    x = rng(n).uniform(-2, 2, size=n).astype(np.float32)
    if n >= 5:
        x[:5] = [0.5, 1.5, 2.5, -0.5, -1.5]   # half-way cases: round half to even
    y = cpu(ops.round_fwd(gpu(x)))
    np.testing.assert_array_equal(y, np.round(x))


def test_round_full_size_property(ops):
    # BASELINE config 2 size: [32,224,224,8]; idempotence + values in {0,1} on softmax outputs
    x = torch.rand(32 * 224 * 224 * 8, device="cuda")
    y = ops.round_fwd(x)
    yy = ops.round_fwd(y)
    assert torch.equal(y, yy)
    assert set(torch.unique(y).tolist()) <= {0.0, 1.0}
    assert torch.equal(y, torch.round(x))


@pytest.mark.parametrize("C", [5, 8])
def test_softmax_round(ops, C):
    x = rng(C).normal(0, 3, size=(2, 17, 19, C)).astype(np.float32)
    p, r = ops.softmax_fwd(gpu(x), want_round=True)
    pr = R.softmax(t(x)).numpy()
    assert rel_l2(cpu(p), pr) < 1e-6
    # bit-exact given the same pre-activation (the kernel's own probabilities)
    np.testing.assert_array_equal(cpu(r), np.round(cpu(p)))
    assert cpu(r).sum(-1).max() <= 1.0   # at most one channel above 0.5
    dp = rng(1).normal(size=x.shape).astype(np.float32)
    xt = t(x, grad=True)
    (R.softmax(xt) * t(dp)).sum().backward()
    dx = ops.softmax_bwd(p, gpu(dp))
    assert rel_l2(cpu(dx), xt.grad.numpy()) < 1e-5


# ------------------------------------------------------------------ activations
@pytest.mark.parametrize("act,alpha", [(1, 0.0), (2, 0.3), (2, 0.2), (3, 0.0)])
def test_activations(ops, act, alpha):
    x = rng(3).normal(size=(3, 9, 7, 8)).astype(np.float32)
    x.ravel()[:7] = 0.0   # exact zeros: derivative must be 0 for relu / leaky relu
    g = rng(4).normal(size=x.shape).astype(np.float32)
    xt = t(x, grad=True)
    yr = {1: R.relu, 2: lambda v: R.leaky_relu(v, alpha), 3: torch.tanh}[act](xt)
    (yr * t(g)).sum().backward()
    y = ops.act_fwd(gpu(x), act, alpha)
    assert rel_l2(cpu(y), yr.detach().numpy()) < 1e-6
    dx = ops.act_bwd(gpu(g), y, act, alpha)
    assert rel_l2(cpu(dx), xt.grad.numpy()) < 1e-5
    if act in (1, 2):
        assert np.all(cpu(dx).ravel()[:7] == 0.0)


def test_add_max_axpby(ops):
    a = rng(0).normal(size=1001).astype(np.float32)
    b = rng(1).normal(size=1001).astype(np.float32)
    b[:100] = a[:100]   # ties
    g = rng(2).normal(size=1001).astype(np.float32)
    np.testing.assert_array_equal(cpu(ops.add(gpu(a), gpu(b))), a + b)
    np.testing.assert_array_equal(cpu(ops.max_fwd(gpu(a), gpu(b))), np.maximum(a, b))
    da, db = ops.max_bwd(gpu(a), gpu(b), gpu(g))
    at, bt = t(a, grad=True), t(b, grad=True)
    (R.tf_maximum(at, bt) * t(g)).sum().backward()
    np.testing.assert_array_equal(cpu(da), at.grad.numpy())
    np.testing.assert_array_equal(cpu(db), bt.grad.numpy())
    assert np.all(cpu(db)[:100] == 0)   # ties go to the first input
    y = gpu(b)
    ops.axpby_(2.0, gpu(a), -1.0, y)
    assert rel_l2(cpu(y), 2 * a - b) < 1e-6


def test_cast_copy_gather(ops):
    x = rng(0).normal(size=(2, 5, 6, 8)).astype(np.float32)
    xb = ops.cast(gpu(x), torch.bfloat16)
    np.testing.assert_array_equal(cpu(xb), t(x).to(torch.bfloat16).float().numpy())
    np.testing.assert_array_equal(cpu(ops.cast(xb, torch.float32)), cpu(xb))
    y = rng(1).normal(size=(2, 5, 6, 1)).astype(np.float32)
    cat = ops.concat_channels([gpu(x), gpu(y)])
    np.testing.assert_array_equal(cpu(cat), np.concatenate([x, y], -1))
    np.testing.assert_array_equal(cpu(ops.slice_channels(cat, 2, 4)), x[..., 2:6])
    idx = np.array([3, 0, 2], np.int32)
    src = rng(2).normal(size=(4, 3, 3, 4)).astype(np.float32)
    np.testing.assert_array_equal(cpu(ops.gather_rows(gpu(src), torch.as_tensor(idx).cuda())), src[idx])


# ------------------------------------------------------------------ FiLM
def test_film(ops):
    B, H, W, C = 3, 20, 24, 8
    x = rng(0).normal(size=(B, H, W, C)).astype(np.float32)
    gm = rng(1).normal(size=(B, C)).astype(np.float32)
    bt = rng(2).normal(size=(B, C)).astype(np.float32)
    g = rng(3).normal(size=x.shape).astype(np.float32)
    xt, gt, btt = t(x, grad=True), t(gm, grad=True), t(bt, grad=True)
    yr = R.film(xt, gt, btt)
    (yr * t(g)).sum().backward()
    y = ops.film_fwd(gpu(x), gpu(gm), gpu(bt))
    assert rel_l2(cpu(y), yr.detach().numpy()) < 1e-6
    dx, dg, db = ops.film_bwd(gpu(g), gpu(x), gpu(gm))
    assert rel_l2(cpu(dx), xt.grad.numpy()) < 1e-6
    assert rel_l2(cpu(dg), gt.grad.numpy()) < 1e-5
    assert rel_l2(cpu(db), btt.grad.numpy()) < 1e-5
    # KAT: FiLM(gamma=1, beta=0) is the identity (layers/film.py:36)
    one, zero = gpu(np.ones((B, C), np.float32)), gpu(np.zeros((B, C), np.float32))
    np.testing.assert_array_equal(cpu(ops.film_fwd(gpu(x), one, zero)), x)


def test_film_block_tail_fused(ops):
    """l1 + LeakyReLU(0.3)(FiLM(l2, gamma, beta)) (model_components/decoder.py:50-54) as one kernel each way:
    bit-identical to the three separate kernels forward, and the oracle's gradients backward"""
    from multimodal_segmentation_b200._lib import ACT_LRELU
    B, H, W, C = 3, 20, 24, 8
    x = rng(0).normal(size=(B, H, W, C)).astype(np.float32)
    res = rng(4).normal(size=(B, H, W, C)).astype(np.float32)
    gm = rng(1).normal(size=(B, C)).astype(np.float32)
    bt = rng(2).normal(size=(B, C)).astype(np.float32)
    g = rng(3).normal(size=x.shape).astype(np.float32)
    xt, rt, gt, btt = t(x, grad=True), t(res, grad=True), t(gm, grad=True), t(bt, grad=True)
    yr = rt + R.leaky_relu(R.film(xt, gt, btt), 0.3)
    (yr * t(g)).sum().backward()
    y = ops.film_act_add_fwd(gpu(x), gpu(gm), gpu(bt), gpu(res), ACT_LRELU, 0.3)
    sep = ops.add(gpu(res), ops.act_fwd(ops.film_fwd(gpu(x), gpu(gm), gpu(bt)), ACT_LRELU, 0.3))
    assert torch.equal(y, sep)
    assert rel_l2(cpu(y), yr.detach().numpy()) < 1e-6
    dx, dg, db = ops.film_act_add_bwd(gpu(g), gpu(x), gpu(gm), gpu(bt), ACT_LRELU, 0.3)
    assert rel_l2(cpu(dx), xt.grad.numpy()) < 1e-6
    assert rel_l2(cpu(dg), gt.grad.numpy()) < 1e-5
    assert rel_l2(cpu(db), btt.grad.numpy()) < 1e-5
    assert np.array_equal(rt.grad.numpy(), g)            # the residual branch receives dy unchanged


# ------------------------------------------------------------------ batch norm
@pytest.mark.parametrize("C,act", [(64, 1), (128, 0), (1024, 1)])
def test_batchnorm_train(ops, C, act):
    N, H, W = 2, 12, 10
    x = (rng(0).normal(size=(N, H, W, C)) * 2 + 0.5).astype(np.float32)
    gm = rng(1).uniform(0.5, 1.5, size=C).astype(np.float32)
    bt = rng(2).normal(size=C).astype(np.float32)
    g = rng(3).normal(size=x.shape).astype(np.float32)
    mm, mv = np.zeros(C, np.float32), np.ones(C, np.float32)
    xt, gt, btt = t(x, grad=True), t(gm, grad=True), t(bt, grad=True)
    yr, mean_r, var_r = R.batchnorm_train(xt, gt, btt)
    if act:
        yr = R.relu(yr)
    (yr * t(g)).sum().backward()
    dmm, dmv = gpu(mm), gpu(mv)
    mean, rstd = ops.bn_stats_finalize(gpu(x), 1e-3, 0.99, dmm, dmv)
    assert rel_l2(cpu(mean), mean_r.detach().numpy()) < 1e-5
    assert rel_l2(cpu(rstd), torch.rsqrt(var_r + 1e-3).detach().numpy()) < 1e-5
    emm, emv = R.bn_moving_update(t(mm), t(mv), mean_r.detach(), var_r.detach(), N * H * W)
    assert rel_l2(cpu(dmm), emm.numpy()) < 1e-5 and rel_l2(cpu(dmv), emv.numpy()) < 1e-5
    y = ops.bn_apply(gpu(x), mean, rstd, gpu(gm), gpu(bt), act)
    assert rel_l2(cpu(y), yr.detach().numpy()) < 1e-5
    yb = ops.bn_apply(gpu(x), mean, rstd, gpu(gm), gpu(bt), act, out_dtype=torch.bfloat16)
    assert rel_l2(cpu(yb), yr.detach().numpy()) < 5e-3   # bf16 storage: 2^-9 relative
    dgm, dbt = ops.zeros(C), ops.zeros(C)
    dx = ops.bn_bwd(gpu(g), gpu(x), mean, rstd, gpu(gm), gpu(bt), act, dgm, dbt)
    assert rel_l2(cpu(dx), xt.grad.numpy()) < FP32_TOL
    assert rel_l2(cpu(dgm), gt.grad.numpy()) < FP32_TOL
    assert rel_l2(cpu(dbt), btt.grad.numpy()) < FP32_TOL


def test_batchnorm_infer(ops):
    C = 64
    x = rng(0).normal(size=(2, 6, 6, C)).astype(np.float32)
    gm, bt = rng(1).uniform(.5, 1.5, C).astype(np.float32), rng(2).normal(size=C).astype(np.float32)
    mm, mv = rng(3).normal(size=C).astype(np.float32), rng(4).uniform(.5, 2, C).astype(np.float32)
    g = rng(5).normal(size=x.shape).astype(np.float32)
    xt = t(x, grad=True)
    yr = R.relu(R.batchnorm_infer(xt, t(gm), t(bt), t(mm), t(mv)))
    (yr * t(g)).sum().backward()
    rstd = ops.bn_rstd_from_var(gpu(mv), 1e-3)
    y = ops.bn_apply(gpu(x), gpu(mm), rstd, gpu(gm), gpu(bt), 1)
    assert rel_l2(cpu(y), yr.detach().numpy()) < 1e-5
    dx = ops.bn_bwd_frozen(gpu(g), gpu(x), gpu(mm), rstd, gpu(gm), gpu(bt), 1)
    assert rel_l2(cpu(dx), xt.grad.numpy()) < 1e-5


# ------------------------------------------------------------------ pooling / resampling
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H,W", [(8, 12), (9, 7)])
def test_maxpool_upsample(ops, dtype, H, W):
    C = 8
    x = np.round(rng(0).normal(size=(2, H, W, C)) * 2).astype(np.float32)   # many ties
    g = rng(1).normal(size=(2, H // 2, W // 2, C)).astype(np.float32)
    g = t(g).to(dtype).float().numpy()
    xt = t(x, grad=True)
    yr = R.maxpool2(xt)
    (yr * t(g)).sum().backward()
    xd = gpu(x, dtype)
    y = ops.maxpool2_fwd(xd)
    np.testing.assert_array_equal(cpu(y), yr.detach().numpy())
    dx = ops.maxpool2_bwd(xd, gpu(g, dtype))
    np.testing.assert_array_equal(cpu(dx), xt.grad.numpy())   # first max in row-major order, bit-exact
    u = ops.upsample2_fwd(xd)
    np.testing.assert_array_equal(cpu(u), R.upsample2(t(x)).numpy())
    gu = rng(2).normal(size=(2, 2 * H, 2 * W, C)).astype(np.float32)
    xt2 = t(x, grad=True)
    (R.upsample2(xt2) * t(gu)).sum().backward()
    du = ops.upsample2_bwd(gpu(gu, dtype))
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert rel_l2(cpu(du), xt2.grad.numpy()) < tol


def test_resize_nn(ops):
    x = rng(0).normal(size=(2, 32, 32, 8)).astype(np.float32)
    for ho in (1, 2, 4, 8, 16, 32):
        y = ops.resize_nn_fwd(gpu(x), ho, ho)
        np.testing.assert_array_equal(cpu(y), R.resize_nn(t(x), ho, ho).numpy())
        g = rng(ho).normal(size=(2, ho, ho, 8)).astype(np.float32)
        xt = t(x, grad=True)
        (R.resize_nn(xt, ho, ho) * t(g)).sum().backward()
        np.testing.assert_array_equal(cpu(ops.resize_nn_bwd(gpu(g), 32, 32)), xt.grad.numpy())


# ------------------------------------------------------------------ general convolution
CONV_CASES = [
    # N,H,W,Cin,Cout,k,stride,pad      (reference call sites)
    (2, 12, 10, 1, 64, 3, 1, 1),     # UNet first conv (models/unet.py:95)
    (2, 12, 10, 8, 64, 3, 1, 1),     # segmentor conv1
    (2, 12, 10, 64, 8, 1, 1, 0),     # conv_anatomy 1x1
    (2, 12, 10, 64, 5, 1, 1, 0),     # segmentor head
    (2, 12, 10, 8, 8, 3, 1, 1),      # FiLM decoder
    (2, 12, 10, 8, 1, 1, 1, 0),      # decoder head
    (2, 17, 15, 9, 16, 3, 2, 0),     # modality encoder, stride 2 valid
    (2, 20, 18, 16, 20, 5, 1, 0),    # locnet 5x5 valid
    (2, 18, 16, 4, 64, 4, 2, 0),     # discriminator first layer
    (1, 13, 11, 64, 128, 4, 2, 0),   # discriminator block
    (1, 9, 9, 32, 48, 4, 1, 0),      # discriminator last block (stride 1)
    (2, 8, 8, 64, 64, 3, 1, 1),      # wide 3x3 (also served by the tcgen05 path)
    (2, 30, 26, 20, 20, 5, 1, 0),    # locnet conv2/3
    (2, 21, 19, 16, 32, 3, 2, 0),    # modality encoder conv2
    (2, 15, 13, 32, 64, 3, 2, 0),    # modality encoder conv3
    (3, 18, 16, 1, 64, 4, 2, 0),     # image discriminator first layer
    (2, 12, 10, 16, 1, 1, 1, 0),     # SPADE decoder head
]


@pytest.mark.parametrize("small", [True, False])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_generic(ops, case, small):
    """small=True: the direct narrow-layer kernels where they apply; False: the general tiled path"""
    ops.USE_SMALL = small
    N, H, W, Cin, Cout, k, s, p = case
    r = rng(sum(case))
    x = r.normal(size=(N, H, W, Cin)).astype(np.float32)
    w = (r.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = r.normal(size=Cout).astype(np.float32)
    xt, wt, bt = t(x, grad=True), t(w, grad=True), t(b, grad=True)
    yr = R.conv2d(xt, wt, bt, stride=s, padding="same" if p else "valid")
    g = r.normal(size=tuple(yr.shape)).astype(np.float32)
    (yr * t(g)).sum().backward()
    y = ops.conv2d_fwd(gpu(x), gpu(w), gpu(b), s, p)
    assert tuple(y.shape) == tuple(yr.shape)
    assert rel_l2(cpu(y), yr.detach().numpy()) < 1e-5
    # second, independent oracle statement
    assert rel_l2(cpu(y), R.conv2d_loops(x, w, b, s, p)) < 1e-5
    dx = ops.conv2d_dgrad(gpu(g), gpu(w), x.shape, s, p)
    assert rel_l2(cpu(dx), xt.grad.numpy()) < 1e-5
    dw, db = ops.zeros(*w.shape), ops.zeros(Cout)
    ops.conv2d_wgrad(gpu(x), gpu(g), dw, db, s, p)
    assert rel_l2(cpu(dw), wt.grad.numpy()) < FP32_TOL
    assert rel_l2(cpu(db), bt.grad.numpy()) < FP32_TOL
    ops.USE_SMALL = True


def test_conv_fused_activation(ops):
    r = rng(7)
    x = r.normal(size=(1, 6, 6, 8)).astype(np.float32)
    w = r.normal(size=(3, 3, 8, 8)).astype(np.float32) * 0.2
    b = r.normal(size=8).astype(np.float32)
    y = ops.conv2d_fwd(gpu(x), gpu(w), gpu(b), 1, 1, act=2, alpha=0.3)
    yr = R.leaky_relu(R.conv2d(t(x), t(w), t(b), 1, "same"), 0.3)
    assert rel_l2(cpu(y), yr.numpy()) < 1e-5


# ------------------------------------------------------------------ dense
@pytest.mark.parametrize("B,K,N", [(4, 1000, 32), (32, 21632, 32), (3, 4802, 100), (5, 270848 // 8, 1), (2, 8, 300), (32, 8, 6272), (192, 8, 6272),
                                   (32, 270848, 1), (3, 1000, 2), (7, 4004, 4), (4, 1002, 1), (33, 512, 3),
                                   (32, 48020, 100), (5, 1028, 8), (40, 4100, 52), (32, 2052, 64), (9, 1500, 100)])
def test_dense(ops, B, K, N):
    r = rng(B + N)
    x = r.normal(size=(B, K)).astype(np.float32)
    w = (r.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
    b = r.normal(size=N).astype(np.float32)
    g = r.normal(size=(B, N)).astype(np.float32)
    xt, wt, bt = t(x, grad=True), t(w, grad=True), t(b, grad=True)
    yr = R.dense(xt, wt, bt)
    (yr * t(g)).sum().backward()
    assert rel_l2(cpu(ops.dense_fwd(gpu(x), gpu(w), gpu(b))), yr.detach().numpy()) < 1e-5
    assert rel_l2(cpu(ops.dense_bwd_data(gpu(g), gpu(w))), xt.grad.numpy()) < 1e-5
    dw, db = ops.zeros(K, N), ops.zeros(N)
    ops.dense_bwd_weight(gpu(x), gpu(g), dw, db)
    assert rel_l2(cpu(dw), wt.grad.numpy()) < 1e-5
    assert rel_l2(cpu(db), bt.grad.numpy()) < 1e-5


# ------------------------------------------------------------------ thin plate spline
def test_tps_general_solve_and_apply(ops):
    r = rng(0)
    B, n, k, m = 3, 25, 2, 500
    c = R.nDgrid((5, 5)).repeat(B, 1, 1).numpy() + r.normal(size=(B, n, 2)).astype(np.float32) * 0.02
    f = r.normal(size=(B, n, k)).astype(np.float32)
    q = r.uniform(0, 1, size=(B, m, 2)).astype(np.float32)
    w, v = ops.tps_solve(gpu(c), gpu(f))
    wr, vr = R.solve_interpolation(t(c, torch.float64), t(f, torch.float64), 2)
    assert rel_l2(cpu(w), wr.numpy()) < 2e-3     # fp32 LU of a cond~4e2 system
    out = ops.tps_apply(gpu(q), gpu(c), w, v)
    outr = R.interpolate_spline(t(c, torch.float64), t(f, torch.float64), t(q, torch.float64), 2)
    assert rel_l2(cpu(out), outr.numpy()) < 1e-3
    # KAT: the spline reproduces train_values at train_points (interpolate_spline.py:225-227)
    at_c = ops.tps_apply(gpu(c), gpu(c), w, v)
    assert np.abs(cpu(at_c) - f).max() < 2e-3


@pytest.mark.parametrize("H,W", [(16, 16), (24, 40)])
def test_tps_warp_forward_backward(ops, H, W):
    r = rng(H)
    B, C = 3, 8
    vol = r.uniform(size=(B, H, W, C)).astype(np.float32)
    theta = (r.normal(size=(B, 25, 2)) * 0.03).astype(np.float32)
    theta[0] = 0.0
    theta[2, :, 1] += 0.6   # push part of the sampling grid outside the image (zero padding path)
    g = r.normal(size=vol.shape).astype(np.float32)
    vt, tt = t(vol, torch.float64, grad=True), t(theta, torch.float64, grad=True)
    outr, locr = R.thin_plate_spline_2d(vt, tt)
    (outr * t(g, torch.float64)).sum().backward()
    out, locs = ops.tps_warp_fwd(gpu(vol), gpu(theta), want_locs=True)
    assert np.abs(cpu(locs) - locr.detach().numpy()).max() < 2e-3      # pixels
    assert rel_l2(cpu(out), outr.detach().numpy()) < 2e-4
    # KAT: theta = 0 is the identity warp (locnet head is zero-initialised, stn_spline.py:116)
    assert np.abs(cpu(out)[0] - vol[0]).max() < 1e-4
    dvol, dtheta = ops.tps_warp_bwd(gpu(vol), gpu(theta), gpu(g))
    assert rel_l2(cpu(dvol)[1:], vt.grad.numpy()[1:]) < 2e-4
    # The coordinate gradient of bilinear sampling is one-sided at integer pixel positions.  With
    # theta == 0 (sample 0) every location sits exactly on the lattice, so a 1e-6 px rounding
    # difference picks the other cell: the comparison is only meaningful for generic theta.
    assert rel_l2(cpu(dtheta)[1:], tt.grad.numpy()[1:]) < 2e-3


def test_resampler_matches_grid_sample(ops):
    r = rng(1)
    B, H, W, C, m = 2, 9, 11, 4, 300
    vol = r.normal(size=(B, H, W, C)).astype(np.float32)
    warp = np.stack([r.uniform(-2, W + 1, size=(B, m)), r.uniform(-2, H + 1, size=(B, m))], -1).astype(np.float32)
    out = ops.resampler_fwd(gpu(vol), gpu(warp))
    assert rel_l2(cpu(out), R.resampler(t(vol), t(warp)).numpy()) < 1e-5
    # independent cross-oracle: torch grid_sample(bilinear, zeros, align_corners=True)
    gx = 2 * warp[..., 0] / (W - 1) - 1
    gy = 2 * warp[..., 1] / (H - 1) - 1
    grid = torch.as_tensor(np.stack([gx, gy], -1))[:, :, None, :]
    gs = torch.nn.functional.grid_sample(t(vol).permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="zeros",
                                         align_corners=True)
    gs = gs[:, :, :, 0].permute(0, 2, 1).numpy()
    assert np.abs(cpu(out) - gs).max() < 1e-4


# ------------------------------------------------------------------ losses
@pytest.mark.parametrize("use_bce,Ct", [(1, 5), (0, 4), (0, 5)])
def test_segloss(ops, use_bce, Ct):
    r = rng(use_bce + Ct)
    B, H, W, Cp = 3, 14, 12, 5
    logits = r.normal(size=(B, H, W, Cp)).astype(np.float32)
    pred = R.softmax(t(logits)).numpy()
    lab = r.randint(0, 5, size=(B, H, W))
    tgt = np.eye(5, dtype=np.float32)[lab][..., :Ct]
    pt = t(pred, grad=True)
    if use_bce:
        lr = 10.0 * R.combined_dice_bce(t(tgt), pt, 4)
    else:
        lr = 10.0 * R.dice_loss(t(tgt), pt, 4)
    lr.backward()
    loss = ops.zeros(1)
    dpred = ops.segloss(gpu(pred), gpu(tgt), 4, use_bce, 10.0, loss)
    assert abs(cpu(loss)[0] - lr.item()) < 1e-4 * abs(lr.item())
    assert rel_l2(cpu(dpred), pt.grad.numpy()) < FP32_TOL
    # KAT: dice(x, x) on a one-hot mask gives (almost) zero loss (costs.py:43-48)
    l0 = ops.zeros(1)
    ops.segloss(gpu(tgt), gpu(tgt), min(4, Ct), 0, 1.0, l0, want_grad=False)
    assert abs(cpu(l0)[0]) < 1e-6


def test_l1l2_vae(ops):
    r = rng(0)
    p = r.normal(size=(4, 10, 10, 1)).astype(np.float32)
    q = r.normal(size=p.shape).astype(np.float32)
    for kind, fn in ((0, R.mae), (1, R.mse)):
        pt = t(p, grad=True)
        lr = 3.0 * fn(t(q), pt)
        lr.backward()
        loss = ops.zeros(1)
        dp = ops.l1l2_loss(gpu(p), gpu(q), kind, 3.0, loss)
        assert abs(cpu(loss)[0] - lr.item()) < 1e-5 * abs(lr.item())
        assert rel_l2(cpu(dp), pt.grad.numpy()) < 1e-6
    d = r.normal(size=(6, 1)).astype(np.float32)
    dt_ = t(d, grad=True)
    lr = R.mse(torch.ones(6, 1), dt_)
    lr.backward()
    loss = ops.zeros(1)
    dd = ops.l1l2_loss(gpu(d), None, 1, 1.0, loss, cval=1.0)
    assert abs(cpu(loss)[0] - lr.item()) < 1e-6 and rel_l2(cpu(dd), dt_.grad.numpy()) < 1e-6
    mu, lv, eps = (r.normal(size=(5, 8)).astype(np.float32) for _ in range(3))
    gz = r.normal(size=(5, 8)).astype(np.float32)
    mt, lt = t(mu, grad=True), t(lv, grad=True)
    zr = R.sampling(mt, lt, t(eps))
    klr = R.kl(mt, lt)
    ((zr * t(gz)).sum() + 0.1 * klr.mean()).backward()
    loss = ops.zeros(1)
    z, klv = ops.vae_fwd(gpu(mu), gpu(lv), gpu(eps), 0.1, loss)
    assert rel_l2(cpu(z), zr.detach().numpy()) < 1e-6 and rel_l2(cpu(klv), klr.detach().numpy()) < 1e-5
    assert abs(cpu(loss)[0] - 0.1 * klr.mean().item()) < 1e-5
    dmu, dlv = ops.vae_bwd(gpu(mu), gpu(lv), gpu(eps), gpu(gz), 0.1)
    assert rel_l2(cpu(dmu), mt.grad.numpy()) < 1e-5 and rel_l2(cpu(dlv), lt.grad.numpy()) < 1e-5
    # KAT: KL(0,0) = 0
    l0 = ops.zeros(1)
    _, k0 = ops.vae_fwd(ops.zeros(2, 8), ops.zeros(2, 8), ops.zeros(2, 8), 1.0, l0)
    assert np.all(cpu(k0) == 0)


def test_spectral_reg(ops):
    r = rng(0)
    dim, cout = 4 * 4 * 16, 32
    W = (r.normal(size=(4, 4, 16, cout)) * 0.1).astype(np.float32)
    u0 = r.uniform(-1, 1, size=(dim, 1)).astype(np.float32)
    wt = t(W, grad=True)
    lr = R.spectral_reg(wt, t(u0), 10.0)
    lr.backward()
    loss, dW = ops.zeros(1), ops.zeros(dim, cout)
    ops.spectral_reg(gpu(W).view(dim, cout), gpu(u0), 10.0, loss, dW)
    assert abs(cpu(loss)[0] - lr.item()) < 1e-4 * abs(lr.item())
    assert rel_l2(cpu(dW), wt.grad.numpy().reshape(dim, cout)) < 1e-3   # sign() flips only at |d|~0


def test_adam(ops):
    r = rng(0)
    n = 10007
    p, g = r.normal(size=n).astype(np.float32), r.normal(size=n).astype(np.float32)
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    dp, dm, dv = gpu(p), gpu(m), gpu(v)
    shadow = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    pr, mr, vr = p.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for step in (1, 2, 3):
        import math
        lr_t = 1e-4 * math.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        ops.adam_step(dp, gpu(g), dm, dv, shadow, lr_t)
        pr, mr, vr = R.adam_step(pr, g.astype(np.float64), mr, vr, step)
    assert rel_l2(cpu(dp), pr) < 1e-6 and rel_l2(cpu(dm), mr) < 1e-6 and rel_l2(cpu(dv), vr) < 1e-6
    np.testing.assert_array_equal(cpu(shadow), cpu(dp.to(torch.bfloat16)))


# ------------------------------------------------------------------ SPADE / instance norm / balancer
def test_spade(ops):
    r = rng(0)
    B, H, W, C = 2, 8, 8, 16
    x = (r.normal(size=(B, H, W, C)) * 3 + 1).astype(np.float32)
    gm = r.normal(size=x.shape).astype(np.float32) * 0.5
    bt = r.normal(size=x.shape).astype(np.float32) * 0.5
    g = r.normal(size=x.shape).astype(np.float32)
    xt, gt, btt = t(x, torch.float64, grad=True), t(gm, torch.float64, grad=True), t(bt, torch.float64, grad=True)
    yr = R.leaky_relu(R.spade_cond(R.instance_norm_axis_none(xt), gt, btt), 0.2)
    (yr * t(g, torch.float64)).sum().backward()
    acc = ops.in_stats(gpu(x))
    y = ops.spade_fwd(gpu(x), acc, gpu(gm), gpu(bt))
    assert rel_l2(cpu(y), yr.detach().numpy()) < 1e-5
    dx, dg, db = ops.spade_bwd(gpu(g), gpu(x), acc, gpu(gm), gpu(bt))
    assert rel_l2(cpu(dg), gt.grad.numpy()) < 1e-5
    assert rel_l2(cpu(db), btt.grad.numpy()) < 1e-5
    assert rel_l2(cpu(dx), xt.grad.numpy()) < FP32_TOL
    # KAT: SPADE_COND(gamma=0, beta=0) is the identity on the normalised input (spade.py:55)
    z = ops.zeros(*x.shape)
    y0 = ops.spade_fwd(gpu(x), acc, z, z, act=0)
    assert rel_l2(cpu(y0), R.instance_norm_axis_none(t(x)).numpy()) < 1e-5


def test_pair_dice(ops):
    r = rng(0)
    a = (r.uniform(size=(3, 10, 10, 8)) > 0.5).astype(np.float32)
    b = (r.uniform(size=(3, 10, 10, 8)) > 0.5).astype(np.float32)
    assert rel_l2(cpu(ops.pair_dice(gpu(a), gpu(b))), R.pair_dice(t(a), t(b)).numpy()) < 1e-6


# ------------------------------------------------------------------ error behaviour of the boundary
def test_error_codes(ops):
    from multimodal_segmentation_b200 import _lib
    x = torch.zeros(16, device="cuda")
    with pytest.raises(_lib.DafkError):
        _lib.call("round_fwd", x.data_ptr() + 4, x, 8, None)      # misaligned
    with pytest.raises(_lib.DafkError):
        _lib.call("softmax_fwd", x, x, None, 2, 7, None)          # unsupported channel count
    with pytest.raises(_lib.DafkError):
        ops.round_fwd(torch.zeros(4))                             # CPU tensor: no fallback


# ------------------------------------------------------------------ augmentation
def _keras_rotate(x, theta):
    """keras 2.1.6 ImageDataGenerator.random_transform with only a rotation + apply_transform(fill_mode='nearest'),
    restated on scipy: x [H,W,C], theta in radians"""
    from scipy import ndimage as ndi
    h, w = x.shape[0], x.shape[1]
    rot = np.array([[np.cos(theta), -np.sin(theta), 0], [np.sin(theta), np.cos(theta), 0], [0, 0, 1]])
    o_x, o_y = float(h) / 2 + 0.5, float(w) / 2 + 0.5
    offset = np.array([[1, 0, o_x], [0, 1, o_y], [0, 0, 1]])
    reset = np.array([[1, 0, -o_x], [0, 1, -o_y], [0, 0, 1]])
    m = offset @ rot @ reset
    chans = [ndi.affine_transform(x[..., c], m[:2, :2], m[:2, 2], order=1, mode="nearest", cval=0.0)
             for c in range(x.shape[-1])]
    return np.stack(chans, -1)


@pytest.mark.parametrize("shape", [(3, 64, 64, 1), (2, 37, 53, 5), (2, 48, 40, 3)])
def test_rotation_augmentation_matches_scipy(ops, shape):
    r = rng(7)
    x = r.uniform(-1, 1, size=shape).astype(np.float32)
    theta = np.deg2rad(r.uniform(-20, 20, size=shape[0])).astype(np.float32)
    theta[0] = 0.0
    y = cpu(ops.rotate_bilinear(gpu(x), gpu(theta)))
    ref = np.stack([_keras_rotate(x[b].astype(np.float64), float(theta[b])) for b in range(shape[0])], 0)
    assert np.array_equal(y[0], x[0])                 # zero angle is the identity
    assert np.abs(y - ref).max() < 2e-4, np.abs(y - ref).max()     # fp32 coordinates vs fp64: ~1e-5 px * gradient



@pytest.mark.parametrize("M,C,dt", [(1000, 64, "bf16"), (3 * 54 * 54, 128, "bf16"), (777, 256, "f32"), (513, 8, "f32"),
                                    (300, 20, "f32"), (129, 1024, "bf16")])
def test_colsum_accumulates_bias_gradient(ops, M, C, dt):
    """out[c] += sum_rows x[row][c]: the wide (8 channels per thread) kernel for power-of-two C, the generic one else"""
    tdt = torch.bfloat16 if dt == "bf16" else torch.float32
    x = rng(M + C).normal(size=(M, C)).astype(np.float32)
    xd = gpu(x, tdt)
    base = rng(1).normal(size=C).astype(np.float32)
    out = gpu(base)
    ops.colsum_(xd, out)
    ref = base.astype(np.float64) + xd.float().cpu().numpy().astype(np.float64).sum(0)
    assert rel_l2(cpu(out), ref) < 1e-5


def test_tps_phi_table_path_is_bit_identical(ops):
    """the per-geometry phi table (dafk_tps_phi_table + dafk_tps_warp_fwd_tab) holds exactly the values the in-kernel
    evaluation computes, so both forward kernels return the same bits (warped volume and sampling locations)"""
    B, H, W, C = 11, 37, 45, 8
    vol = gpu(rng(0).uniform(size=(B, H, W, C)).astype(np.float32))
    theta = gpu((rng(1).normal(size=(B, 25, 2)) * 0.05).astype(np.float32))
    old = ops.TPS_PHI_TABLE
    try:
        ops.TPS_PHI_TABLE = True
        a, la = ops.tps_warp_fwd(vol, theta, want_locs=True)
        ops.TPS_PHI_TABLE = False
        b, lb = ops.tps_warp_fwd(vol, theta, want_locs=True)
    finally:
        ops.TPS_PHI_TABLE = old
    assert torch.equal(a, b) and torch.equal(la, lb)
    # backward: the table kernel (compile-time 25 control points, phi and pixel coordinates from the table) against the
    # in-kernel evaluation -- same arithmetic; the scatter into the volume uses atomics, so compare to round-off
    dout = gpu(rng(2).normal(size=(B, H, W, C)).astype(np.float32))
    try:
        ops.TPS_PHI_TABLE = True
        dv_a, dt_a = ops.tps_warp_bwd(vol, theta, dout)
        ops.TPS_PHI_TABLE = False
        dv_b, dt_b = ops.tps_warp_bwd(vol, theta, dout)
    finally:
        ops.TPS_PHI_TABLE = old
    assert rel_l2(cpu(dv_a), cpu(dv_b)) < 1e-6 and rel_l2(cpu(dt_a), cpu(dt_b)) < 1e-6
    # other channel counts / control grids take the generic kernels
    vol4 = gpu(rng(3).uniform(size=(3, 20, 22, 4)).astype(np.float32))
    th9 = gpu((rng(4).normal(size=(3, 9, 2)) * 0.05).astype(np.float32))
    try:
        ops.TPS_PHI_TABLE = True
        a4, _ = ops.tps_warp_fwd(vol4, th9, cp=(3, 3))
        g4 = ops.tps_warp_bwd(vol4, th9, a4, cp=(3, 3))
        ops.TPS_PHI_TABLE = False
        b4, _ = ops.tps_warp_fwd(vol4, th9, cp=(3, 3))
        h4 = ops.tps_warp_bwd(vol4, th9, a4, cp=(3, 3))
    finally:
        ops.TPS_PHI_TABLE = old
    assert torch.equal(a4, b4) and rel_l2(cpu(g4[0]), cpu(h4[0])) < 1e-6 and rel_l2(cpu(g4[1]), cpu(h4[1])) < 1e-6


def test_residual_channel_is_rebuilt_after_the_rotation(ops):
    """the reference augments the nm-channel masks and calls add_residual on the AUGMENTED batch
    (model_executors/dafnet_executor.py:493-494, base_executor.py:83-87), so the background channel is 1 wherever the
    bilinearly rotated mask is not exactly 1.  dafk_mask_residual on the staged batch against the reference order of
    operations in numpy, and the executor's stager applies it to the label arrays of its generators."""
    shape = (3, 33, 29, 5)
    r = rng(3)
    m = (r.uniform(size=shape[:3] + (4,)) > 0.6).astype(np.float32)
    theta = r.uniform(-0.3, 0.3, size=3).astype(np.float32)
    staged = np.concatenate([m, np.ones(shape[:3] + (1,), np.float32)], -1)          # residual channel allocated, stale
    rot = ops.rotate_bilinear(gpu(staged), gpu(theta))
    ops.mask_residual_(rot)
    got = cpu(rot)
    rot_masks = np.stack([_keras_rotate(m[b].astype(np.float64), float(theta[b])) for b in range(3)], 0).astype(np.float32)
    want_res = np.ones(shape[:3] + (1,), np.float32)
    for i in range(4):
        want_res[cpu(rot)[..., i:i + 1] == 1] = 0                                   # add_residual on the rotated channels
    assert np.array_equal(got[..., 4:5], want_res)
    assert np.abs(got[..., :4] - rot_masks).max() < 1e-5
    assert 0 < want_res.mean() < 1 and (want_res != (1 - np.clip(rot_masks.sum(-1, keepdims=True), 0, 1))).any()
    # host plumbing: label arrays of a generator are flagged, image arrays are not
    from multimodal_segmentation_b200.model_executors.base_executor import BatchFlow, FlowGroup
    g = FlowGroup([BatchFlow(np.zeros((4, 8, 8, 1), np.float32), 2, 0, 0.0), BatchFlow(np.zeros((4, 8, 8, 5), np.float32), 2, 0, 0.0)], [1])
    assert g.residual_items == (1,)

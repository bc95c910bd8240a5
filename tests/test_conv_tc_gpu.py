"""Parity of the tcgen05/TMEM/TMA convolution path against the oracle.

The kernels take bf16 operands and accumulate in fp32, so the oracle is evaluated on the SAME
bf16-rounded inputs in fp64: what remains is accumulation-order noise (tolerance 1e-4 relative
L2 for forward/dgrad, 1e-3 for the atomically reduced weight gradient).  Against the unrounded
fp32 oracle the bound is the north-star 1e-2 for bf16 conv paths.
"""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from tests.util import cpu, gpu, rel_l2, t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from multimodal_segmentation_b200 import ops as o
    return o


def bf16_round(a):
    return torch.as_tensor(a).to(torch.bfloat16).float().numpy()


CASES = [
    # N, H, W, C0, C1, Cout
    (2, 16, 16, 64, 0, 64),      # exact tiles
    (1, 24, 40, 64, 0, 64),      # partial tiles in both directions
    (2, 16, 16, 64, 64, 64),     # two K sources (concat)
    (2, 16, 16, 128, 0, 128),    # BLOCK_N = 128 path
    (4, 14, 14, 128, 128, 256),  # deep layer shape: multi-image boxes, 2 n-blocks
    (3, 7, 7, 128, 0, 128),      # SPADE head resolution
    (1, 56, 56, 64, 0, 64),
]


@pytest.mark.parametrize("case", CASES)
def test_conv3x3_tc_forward(ops, case):
    N, H, W, C0, C1, Cout = case
    r = np.random.RandomState(sum(case))
    Cin = C0 + C1
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    w = bf16_round((r.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32))
    b = r.normal(size=Cout).astype(np.float32)
    yr = R.conv2d(t(x, torch.float64), t(w, torch.float64), t(b, torch.float64), 1, "same").numpy()
    wp = ops.pack_conv3x3(gpu(w))
    x0 = gpu(x[..., :C0], torch.bfloat16)
    x1 = gpu(x[..., C0:], torch.bfloat16) if C1 else None
    y = ops.conv3x3_tc_fwd(x0, x1, wp, gpu(b), Cout)
    err = rel_l2(cpu(y), yr)
    assert err < 1e-4, err
    yb = ops.conv3x3_tc_fwd(x0, x1, wp, gpu(b), Cout, out_dtype=torch.bfloat16)
    assert rel_l2(cpu(yb), yr) < 5e-3


PAIR_CASES = [
    # N, H, W, C0, C1, Cout      (resident half-weights per CTA: Cout <= 128, (C0 + C1) * Cout * 9 <= 150 KB per half)
    (2, 16, 16, 64, 0, 64),
    (3, 24, 40, 64, 0, 64),      # odd number of tiles: the peer CTA's last tile is empty
    (2, 16, 16, 64, 64, 64),     # two K sources
    (2, 16, 16, 64, 0, 128),
    (2, 16, 16, 128, 0, 128),
    (1, 56, 56, 64, 0, 64),
    (5, 64, 64, 64, 0, 64),      # several tile pairs per cluster: both TMEM buffers and every stage are reused
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv3x3_tc_forward_cta_pairs(ops, case, monkeypatch):
    """tcgen05.mma.cta_group::2 variant of the haloed-tile kernel against the oracle and the single-CTA kernel"""
    N, H, W, C0, C1, Cout = case
    r = np.random.RandomState(sum(case) + 11)
    Cin = C0 + C1
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    w = bf16_round((r.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32))
    b = r.normal(size=Cout).astype(np.float32)
    yr = R.conv2d(t(x, torch.float64), t(w, torch.float64), t(b, torch.float64), 1, "same").numpy()
    wp = ops.pack_conv3x3(gpu(w))
    x0 = gpu(x[..., :C0], torch.bfloat16)
    x1 = gpu(x[..., C0:], torch.bfloat16) if C1 else None
    monkeypatch.setenv("DAFK_CONV_HALO2", "0")
    y0 = ops.conv3x3_tc_fwd(x0, x1, wp, gpu(b), Cout)
    monkeypatch.setenv("DAFK_CONV_HALO2", "1")
    for _ in range(2):
        y1 = ops.conv3x3_tc_fwd(x0, x1, wp, gpu(b), Cout)
        torch.cuda.synchronize()
        assert rel_l2(cpu(y1), yr) < 1e-4
        assert rel_l2(cpu(y1), cpu(y0)) < 1e-6      # same K order per accumulator


@pytest.mark.parametrize("shape", [(3, 64, 64), (3, 64, 128), (3, 80, 72), (4, 4, 64), (3, 1024, 512), (1, 128, 8)])
def test_pack_conv_tiled_transpose_is_bit_identical(ops, shape, monkeypatch):
    k, Cin, Cout = shape
    r = np.random.RandomState(sum(shape))
    w = gpu(r.normal(size=(k, k, Cin, Cout)).astype(np.float32))
    scale = gpu(r.uniform(0.5, 2.0, size=Cout).astype(np.float32))
    monkeypatch.setenv("DAFK_PACK_TILED", "0")
    p0, s0 = ops.pack_conv(w, 0), ops.pack_conv_scaled(w, scale)
    monkeypatch.setenv("DAFK_PACK_TILED", "1")
    p1, s1 = ops.pack_conv(w, 0), ops.pack_conv_scaled(w, scale)
    torch.cuda.synchronize()
    assert torch.equal(p0, p1) and torch.equal(s0, s1)


@pytest.mark.parametrize("case", CASES[:5])
def test_conv3x3_tc_dgrad(ops, case):
    N, H, W, C0, C1, Cout = case
    Cin = C0 + C1
    r = np.random.RandomState(sum(case) + 1)
    w = bf16_round((r.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32))
    dy = bf16_round(r.normal(size=(N, H, W, Cout)).astype(np.float32))
    xt = torch.zeros(N, H, W, Cin, dtype=torch.float64, requires_grad=True)
    (R.conv2d(xt, t(w, torch.float64), None, 1, "same") * t(dy, torch.float64)).sum().backward()
    wpd = ops.pack_conv3x3(gpu(w), for_dgrad=True)     # [9][Cin][Cout], taps mirrored
    dx = ops.conv3x3_tc_fwd(gpu(dy, torch.bfloat16), None, wpd, None, Cin)
    err = rel_l2(cpu(dx), xt.grad.numpy())
    assert err < 1e-4, err


@pytest.mark.parametrize("case", [(2, 16, 16, 64, 64), (1, 24, 40, 64, 64), (2, 16, 16, 128, 128),
                                  (2, 16, 16, 64, 128), (2, 16, 16, 128, 64), (4, 14, 14, 256, 128)])
def test_conv3x3_tc_wgrad(ops, case):
    N, H, W, Cin, Cout = case
    r = np.random.RandomState(sum(case) + 2)
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    dy = bf16_round(r.normal(size=(N, H, W, Cout)).astype(np.float32))
    wt = torch.zeros(3, 3, Cin, Cout, dtype=torch.float64, requires_grad=True)
    (R.conv2d(t(x, torch.float64), wt, None, 1, "same") * t(dy, torch.float64)).sum().backward()
    dw = ops.zeros(3, 3, Cin, Cout)
    ops.conv3x3_tc_wgrad(gpu(x, torch.bfloat16), gpu(dy, torch.bfloat16), dw)
    err = rel_l2(cpu(dw), wt.grad.numpy())
    assert err < 1e-3, err


@pytest.mark.parametrize("case", [(2, 16, 16, 64, 64), (1, 24, 40, 64, 64), (3, 37, 21, 64, 64), (2, 16, 32, 128, 128),
                                  (2, 16, 16, 64, 128), (2, 24, 16, 128, 64), (1, 56, 56, 64, 64), (2, 8, 8, 256, 64)])
def test_conv3x3_tc_wgrad_halo(ops, case):
    """the all-taps-per-CTA weight-gradient kernel (csrc/conv_tc_wgrad_halo.cu): tiles that overhang the image, several
    channel blocks, more tiles than pipeline stages, and accumulation into a non-zero dw at a channel offset"""
    N, H, W, Cin, Cout = case
    r = np.random.RandomState(sum(case) + 5)
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    dy = bf16_round(r.normal(size=(N, H, W, Cout)).astype(np.float32))
    wt = torch.zeros(3, 3, Cin, Cout, dtype=torch.float64, requires_grad=True)
    (R.conv2d(t(x, torch.float64), wt, None, 1, "same") * t(dy, torch.float64)).sum().backward()
    base = r.normal(size=(3, 3, Cin + 64, Cout)).astype(np.float32)
    dw = gpu(base)
    ops.conv3x3_tc_wgrad_halo(gpu(x, torch.bfloat16), gpu(dy, torch.bfloat16), dw, cin_off=64)
    got = cpu(dw) - base
    assert np.array_equal(got[:, :, :64], np.zeros_like(got[:, :, :64]))        # other channel blocks untouched
    err = rel_l2(got[:, :, 64:], wt.grad.numpy())
    assert err < 1e-3, err


def test_conv3x3_tc_wgrad_concat_offset(ops):
    N, H, W, C0, C1, Cout = 2, 16, 16, 64, 64, 64
    r = np.random.RandomState(11)
    x = bf16_round(r.normal(size=(N, H, W, C0 + C1)).astype(np.float32))
    dy = bf16_round(r.normal(size=(N, H, W, Cout)).astype(np.float32))
    wt = torch.zeros(3, 3, C0 + C1, Cout, dtype=torch.float64, requires_grad=True)
    (R.conv2d(t(x, torch.float64), wt, None, 1, "same") * t(dy, torch.float64)).sum().backward()
    dw = ops.zeros(3, 3, C0 + C1, Cout)
    ops.conv3x3_tc_wgrad(gpu(x[..., :C0], torch.bfloat16), gpu(dy, torch.bfloat16), dw, cin_off=0)
    ops.conv3x3_tc_wgrad(gpu(x[..., C0:], torch.bfloat16), gpu(dy, torch.bfloat16), dw, cin_off=C0)
    assert rel_l2(cpu(dw), wt.grad.numpy()) < 1e-3


def test_conv3x3_tc_vs_fp32_oracle_tolerance(ops):
    """north-star bound for the bf16 conv path against the unrounded fp32 reference: <= 1e-2"""
    N, H, W, C, Cout = 2, 28, 28, 128, 128
    r = np.random.RandomState(5)
    x = r.normal(size=(N, H, W, C)).astype(np.float32)
    w = (r.normal(size=(3, 3, C, Cout)) / np.sqrt(9 * C)).astype(np.float32)
    yr = R.conv2d(t(x), t(w), None, 1, "same").numpy()
    y = ops.conv3x3_tc_fwd(ops.cast(gpu(x), torch.bfloat16), None, ops.pack_conv3x3(gpu(w)), None, Cout)
    assert rel_l2(cpu(y), yr) < 1e-2


# ------------------------------------------------------------------ general form: discriminator layers
DISC_CASES = [
    # N, H, W, Cin, Cout, k, stride     (models/discriminator.py:24,39 -- valid padding)
    (2, 111, 111, 64, 128, 4, 2),
    (2, 54, 54, 128, 256, 4, 2),
    (2, 26, 26, 256, 512, 4, 1),
    (3, 31, 29, 64, 64, 4, 2),      # odd sizes: uncovered last row / column in the data gradient
    (2, 20, 22, 64, 128, 5, 1),     # another stride-1 valid shape
    # channel counts that are not multiples of 64 (SPADE decoder, layers/spade.py:14-31): partial 64-channel blocks
    # are zero-padded by TMA out-of-bounds fill and by the packed weights
    (2, 20, 22, 128, 32, 3, 1),
    (2, 20, 22, 32, 16, 3, 1),
    (2, 16, 18, 48, 80, 3, 1),
    (2, 28, 28, 128, 16, 3, 1),
]


@pytest.mark.parametrize("case", DISC_CASES)
def test_conv_tc_general_fwd_dgrad_wgrad(ops, case):
    from multimodal_segmentation_b200 import engine as E
    N, H, W, Cin, Cout, k, s = case
    r = np.random.RandomState(sum(case))
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    w = bf16_round((r.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32))
    b = r.normal(size=Cout).astype(np.float32)
    xt, wt = t(x, torch.float64, grad=True), t(w, torch.float64, grad=True)
    yr = R.conv2d(xt, wt, t(b, torch.float64), s, "valid")
    dy = bf16_round(r.normal(size=tuple(yr.shape)).astype(np.float32))
    (yr * t(dy, torch.float64)).sum().backward()
    E.USE_TC = True
    arena = E.Arena(True)
    conv = E.Conv2D(arena, r, "c", Cin, Cout, k, s, "valid")
    arena.to_device()
    conv.kernel.data.copy_(gpu(w))
    conv.bias.data.copy_(gpu(b))
    tape = E.Tape()
    ctx = E.Ctx(tape, True)
    xv = E.Var(gpu(x, torch.bfloat16), True)
    assert conv.tc_eligible([xv])
    y = conv(ctx, xv)
    assert tuple(y.shape) == tuple(yr.shape)
    assert rel_l2(cpu(y.data), yr.detach().numpy()) < 1e-4
    y.grad = gpu(dy, torch.bfloat16)
    tape.backward()
    assert rel_l2(cpu(xv.grad), xt.grad.numpy()) < 5e-3          # dx is stored as bf16
    assert rel_l2(cpu(conv.kernel.grad), wt.grad.numpy()) < 1e-3
    assert rel_l2(cpu(conv.bias.grad), dy.sum((0, 1, 2))) < 1e-3


@pytest.mark.parametrize("case", [(2, 20, 20, 64, 64, 4), (2, 21, 23, 64, 128, 4), (3, 14, 13, 128, 64, 4), (1, 9, 10, 128, 256, 2),
                                  (33, 27, 27, 64, 128, 4)])
def test_conv_tc_stride2_dgrad_all_parity_classes_in_one_launch(ops, case):
    """dafk_conv_tc_dgrad_s2: data gradient of a valid stride-2 convolution with an even kernel (models/discriminator.py:
    24,39) -- odd and even input extents (rows / columns the convolution never read must come out 0), Cin of 64 and 128
    (both tile widths), a batch that does not divide the image box"""
    N, H, W, Cin, Cout, k = case
    r = np.random.RandomState(sum(case) + 3)
    w = bf16_round((r.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32))
    Ho, Wo = (H - k) // 2 + 1, (W - k) // 2 + 1
    dy = bf16_round(r.normal(size=(N, Ho, Wo, Cout)).astype(np.float32))
    xt = torch.zeros(N, H, W, Cin, dtype=torch.float64, requires_grad=True)
    (R.conv2d(xt, t(w, torch.float64), None, 2, "valid") * t(dy, torch.float64)).sum().backward()
    wp4 = ops.pack_conv_s2_all(gpu(w))
    for dt in (torch.float32, torch.bfloat16):
        dx = ops.conv_tc_dgrad_s2(gpu(dy, torch.bfloat16), wp4, (N, H, W, Cin), Cin, k, k, dt)
        assert dx.dtype == dt and tuple(dx.shape) == (N, H, W, Cin)
        assert rel_l2(cpu(dx), xt.grad.numpy()) < (1e-4 if dt == torch.float32 else 4e-3)
    # a source that is a channel window of a wider kernel (second source of a Concatenate)
    if Cin == 128:
        dxw = ops.conv_tc_dgrad_s2(gpu(dy, torch.bfloat16), wp4, (N, H, W, 64), 64, k, k, torch.float32, row_off=64,
                                   rows_per_tap=128)
        assert rel_l2(cpu(dxw), xt.grad.numpy()[..., 64:]) < 1e-4


BN_CASES = [
    # N, H, W, C0, C1, Cout, k, stride     -- every forward kernel: CTA-pair haloed (64 / 128 outputs, resident weights),
    (2, 56, 56, 64, 0, 64, 3, 1),          #    single-CTA haloed, tap-by-tap with one and with several output-channel blocks
    (3, 24, 40, 64, 0, 64, 3, 1),          # partial tiles: dead rows must contribute nothing
    (2, 28, 28, 128, 0, 128, 3, 1),
    (2, 28, 28, 64, 64, 64, 3, 1),         # two sources
    (2, 20, 20, 256, 0, 256, 3, 1),
    (3, 14, 14, 512, 0, 1024, 3, 1),       # four 256-wide output-channel blocks: per-tile flush of the running sums
    (2, 13, 15, 64, 0, 128, 3, 1),
    (1, 9, 9, 1024, 0, 512, 3, 1),
    (5, 64, 64, 64, 0, 64, 3, 1),          # many tiles per CTA: running sums persist across tiles
    (2, 21, 23, 64, 0, 128, 4, 2),         # strided layer (tap-by-tap kernel)
    (2, 16, 16, 64, 0, 96, 3, 1),          # Cout not a multiple of the chunk width
]


@pytest.mark.parametrize("case", BN_CASES)
def test_conv_tc_forward_with_batchnorm_statistics(ops, case, monkeypatch):
    """dafk_conv_tc_fwd_bn: the stored bf16 output is bit-identical to the plain kernel's, and the fp64 accumulators hold
    the per-channel sum / sum of squares of exactly those stored values (what a BatchNormalization after the convolution
    normalises, utils/model_utils.py:10); dafk_bn_finalize on them equals dafk_bn_stats_fused on the map."""
    N, H, W, C0, C1, Cout, k, stride = case
    r = np.random.RandomState(sum(case))
    Cin = C0 + C1
    pad = 1 if (k == 3 and stride == 1) else 0
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    w = bf16_round((r.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32))
    b = r.normal(size=Cout).astype(np.float32)
    wp = ops.pack_conv(gpu(w), 0)
    x0 = gpu(x[..., :C0], torch.bfloat16)
    x1 = gpu(x[..., C0:], torch.bfloat16) if C1 else None
    for halo2 in ("1", "0"):                      # CTA-pair and single-CTA haloed kernels where both apply
        monkeypatch.setenv("DAFK_CONV_HALO2", halo2)
        y_ref = ops.conv_tc_fwd(x0, x1, wp, gpu(b), Cout, k, k, stride, pad, torch.bfloat16)
        y, acc = ops.conv_tc_fwd_bn(x0, x1, wp, gpu(b), Cout, k, k, stride, pad)
        torch.cuda.synchronize()
        assert torch.equal(y, y_ref)
        yd = y.double().reshape(-1, Cout)
        want = torch.cat([yd.sum(0), (yd * yd).sum(0)])
        got = acc.double()
        assert torch.allclose(got, want, rtol=2e-6, atol=1e-4), (got - want).abs().max().item()
    M = y.numel() // Cout
    mean, rstd = ops.bn_finalize_acc(acc, M, 1e-3, 0.99)
    if Cout & (Cout - 1) == 0:                    # the stand-alone statistics kernels take power-of-two channel counts
        mean2, rstd2 = ops.bn_stats_finalize(y, 1e-3, 0.99)
        assert rel_l2(cpu(mean), cpu(mean2)) < 1e-5 and rel_l2(cpu(rstd), cpu(rstd2)) < 1e-5
    yd = y.double().reshape(-1, Cout)
    assert rel_l2(cpu(mean), yd.mean(0).cpu().numpy()) < 1e-5
    assert rel_l2(cpu(rstd), (1.0 / torch.sqrt(yd.var(0, unbiased=False) + 1e-3)).cpu().numpy()) < 1e-5


def test_conv_bn_block_uses_the_fused_statistics(ops, monkeypatch):
    """engine.conv_bn in the training phase: same outputs, moving statistics and gradients with the statistics taken in
    the convolution's epilogue (default) as with the separate statistics pass (DAFK_FUSE_BN_STATS=0), one launch fewer"""
    from multimodal_segmentation_b200 import engine as E

    def run(fuse):
        monkeypatch.setattr(ops, "FUSE_BN_STATS", fuse)
        E.USE_TC = True
        rs = np.random.RandomState(3)
        arena, state = E.Arena(True), E.Arena(False)
        conv = E.Conv2D(arena, rs, "c", 64, 128, 3, 1, "same")       # >= 128 outputs: the layers that use the fused statistics
        bn = E.BatchNorm(arena, state, "n", 128)
        arena.to_device()
        state.to_device()
        x = E.Var(gpu(bf16_round(rs.normal(size=(2, 24, 24, 64)).astype(np.float32)), torch.bfloat16), True)
        tape = E.Tape()
        ctx = E.Ctx(tape, True)
        n0 = ops._lib.launch_count()
        y = E.conv_bn(ctx, conv, bn, x, "relu", torch.bfloat16)
        launches = ops._lib.launch_count() - n0
        y.grad = gpu(bf16_round(rs.normal(size=tuple(y.shape)).astype(np.float32)), torch.bfloat16)
        tape.backward()
        torch.cuda.synchronize()
        return (launches, cpu(y.data), cpu(bn.moving_mean.data), cpu(bn.moving_var.data), cpu(x.grad), cpu(conv.kernel.grad),
                cpu(bn.gamma.grad))

    a, b = run(True), run(False)
    assert a[0] < b[0] + 2            # (memset + finalize) replace the statistics kernel; never more than one extra launch
    assert np.array_equal(a[1], b[1]) or rel_l2(a[1], b[1]) < 1e-2      # bf16 outputs: identical up to last-bit statistics
    for i in range(2, 7):
        assert rel_l2(a[i], b[i]) < 2e-3, (i, rel_l2(a[i], b[i]))

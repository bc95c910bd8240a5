"""Parity of the narrow-channel tcgen05 convolutions (csrc/conv_nc.cu) against the oracle.

Operands are rounded to bf16 by the kernel while it stages them, accumulation is fp32 in tensor
memory; the oracle is evaluated in fp64 on the SAME bf16-rounded operands, so the tolerance
(1e-4 forward / data gradient, 1e-3 for the atomically reduced weight gradient) only covers the
accumulation order.  Against the unrounded fp32 oracle the bound is the north-star 1e-2.
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
    # N, H, W, Cin, Cout, k, pad
    (2, 16, 16, 8, 8, 3, 1),        # FiLM decoder layer (model_components/decoder.py:44-54)
    (3, 37, 45, 8, 8, 3, 1),        # ragged strips / tiles
    (2, 24, 40, 8, 64, 3, 1),       # segmentor conv1 (model_components/segmentor.py:15)
    (2, 24, 24, 1, 64, 3, 1),       # UNet first layer (models/unet.py:95)
    (2, 40, 36, 16, 20, 5, 0),      # locnet conv1 (layers/stn_spline.py:106)
    (2, 30, 30, 20, 20, 5, 0),      # locnet conv2/3
    (2, 20, 28, 8, 1, 1, 0),        # decoder output 1x1 (decoder.py:28)
    (1, 224, 224, 8, 8, 3, 1),      # full-resolution strip geometry
    (2, 12, 12, 9, 16, 3, 1),       # odd channel count (scalar staging path)
    (2, 18, 22, 4, 5, 3, 1),
]


def _mk(case, seed):
    N, H, W, Cin, Cout, k, pad = case
    r = np.random.RandomState(seed)
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    w = bf16_round((r.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32))
    b = r.normal(size=Cout).astype(np.float32)
    return r, x, w, b


def _ref_conv(x, w, b, pad):
    return R.conv2d(t(x, torch.float64), t(w, torch.float64), None if b is None else t(b, torch.float64), 1,
                    "same" if pad else "valid")


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("xdt", ["f32", "bf16"])
def test_conv_nc_forward(ops, case, xdt):
    N, H, W, Cin, Cout, k, pad = case
    r, x, w, b = _mk(case, sum(case))
    yr = _ref_conv(x, w, b, pad).numpy()
    wp = ops.pack_conv_nc(gpu(w), 0)
    xg = gpu(x, torch.float32 if xdt == "f32" else torch.bfloat16)
    y = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad)
    assert tuple(y.shape) == tuple(yr.shape)
    err = rel_l2(cpu(y), yr)
    assert err < 1e-4, err
    # fused LeakyReLU epilogue + bf16 output
    from multimodal_segmentation_b200._lib import ACT_LRELU
    ya = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad, ACT_LRELU, 0.3, torch.bfloat16)
    ref = np.where(yr > 0, yr, 0.3 * yr)
    assert rel_l2(cpu(ya), ref) < 5e-3


@pytest.mark.parametrize("case", CASES[:6] + CASES[8:])
def test_conv_nc_dgrad(ops, case):
    N, H, W, Cin, Cout, k, pad = case
    r, x, w, b = _mk(case, sum(case) + 1)
    xt = torch.zeros(N, H, W, Cin, dtype=torch.float64, requires_grad=True)
    yr = R.conv2d(xt, t(w, torch.float64), None, 1, "same" if pad else "valid")
    dy = bf16_round(r.normal(size=tuple(yr.shape)).astype(np.float32))
    (yr * t(dy, torch.float64)).sum().backward()
    wpd = ops.pack_conv_nc(gpu(w), 1)
    dx = ops.conv_nc_fwd(gpu(dy), wpd, None, Cin, k, k, k - 1 - pad)
    assert tuple(dx.shape) == (N, H, W, Cin)
    err = rel_l2(cpu(dx), xt.grad.numpy())
    assert err < 1e-4, err


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dts", [("f32", "f32"), ("bf16", "bf16")])
def test_conv_nc_wgrad(ops, case, dts):
    N, H, W, Cin, Cout, k, pad = case
    r, x, w, b = _mk(case, sum(case) + 2)
    wt = torch.zeros(k, k, Cin, Cout, dtype=torch.float64, requires_grad=True)
    yr = R.conv2d(t(x, torch.float64), wt, None, 1, "same" if pad else "valid")
    dy = bf16_round(r.normal(size=tuple(yr.shape)).astype(np.float32))
    (yr * t(dy, torch.float64)).sum().backward()
    dw = ops.zeros(k, k, Cin, Cout)
    db = ops.zeros(Cout)
    td = {"f32": torch.float32, "bf16": torch.bfloat16}
    ops.conv_nc_wgrad(gpu(x, td[dts[0]]), gpu(dy, td[dts[1]]), dw, db, pad)
    err = rel_l2(cpu(dw), wt.grad.numpy())
    assert err < 1e-3, err
    assert rel_l2(cpu(db), dy.sum((0, 1, 2))) < 1e-3
    # accumulates into dw (second call doubles it), db optional
    ops.conv_nc_wgrad(gpu(x, td[dts[0]]), gpu(dy, td[dts[1]]), dw, None, pad)
    assert rel_l2(cpu(dw), 2 * wt.grad.numpy()) < 1e-3


def test_conv_nc_vs_fp32_oracle_tolerance(ops):
    """north-star bound for the bf16 conv path against the unrounded fp32 reference: <= 1e-2"""
    r = np.random.RandomState(3)
    x = r.normal(size=(2, 32, 32, 8)).astype(np.float32)
    w = (r.normal(size=(3, 3, 8, 8)) / np.sqrt(72)).astype(np.float32)
    yr = _ref_conv(x, w, None, 1).numpy()
    y = ops.conv_nc_fwd(gpu(x), ops.pack_conv_nc(gpu(w), 0), None, 8, 3, 3, 1)
    assert rel_l2(cpu(y), yr) < 1e-2


# ------------------------------------------------------------------ bf16 input with 8 channels: rows staged by cp.async.bulk
BULK_CASES = [
    # N, H, W, Cin, Cout, k, pad
    (2, 16, 16, 8, 8, 3, 1),
    (3, 37, 45, 8, 8, 3, 1),        # ragged strips / tiles, W not a multiple of 8
    (2, 24, 40, 8, 64, 3, 1),
    (2, 20, 28, 8, 1, 1, 0),
    (2, 30, 30, 8, 20, 5, 0),
    (2, 30, 34, 8, 5, 5, 2),
    (1, 224, 224, 8, 8, 3, 1),
    (64, 160, 48, 8, 8, 3, 1),      # > 4 strips per CTA: every stage is reused, rows outside the image re-zeroed
]


@pytest.mark.parametrize("case", BULK_CASES)
def test_conv_nc_bulk_rows_forward(ops, case, monkeypatch):
    from multimodal_segmentation_b200._lib import ACT_LRELU
    N, H, W, Cin, Cout, k, pad = case
    r, x, w, b = _mk(case, sum(case) + 5)
    yr = _ref_conv(x, w, b, pad).numpy()
    wp = ops.pack_conv_nc(gpu(w), 0)
    xg = gpu(x, torch.bfloat16)
    monkeypatch.setenv("DAFK_NC_BULK", "0")
    y0 = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad)
    a0 = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad, ACT_LRELU, 0.3, torch.bfloat16)
    monkeypatch.setenv("DAFK_NC_BULK", "1")
    for _ in range(2):              # second launch: nothing may depend on leftovers of the first
        y1 = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad)
        a1 = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad, ACT_LRELU, 0.3, torch.bfloat16)
        torch.cuda.synchronize()
        assert rel_l2(cpu(y1), yr) < 1e-4
        assert torch.equal(y0, y1) and torch.equal(a0, a1)      # same operands, same MMA order: bit-identical


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("xdt", ["f32", "bf16"])
def test_conv_nc_twelve_warp_layout(ops, case, xdt, monkeypatch):
    """7 producer warps + MMA issuer + 4 epilogue warps (168-register budget) must give the default kernel's bits"""
    N, H, W, Cin, Cout, k, pad = case
    r, x, w, b = _mk(case, sum(case) + 7)
    wp = ops.pack_conv_nc(gpu(w), 0)
    xg = gpu(x, torch.float32 if xdt == "f32" else torch.bfloat16)
    monkeypatch.setenv("DAFK_NC_L12", "0")
    y0 = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad)
    monkeypatch.setenv("DAFK_NC_L12", "1")
    y1 = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad)
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)


RAW_CASES = CASES + [
    (2, 24, 40, 64, 8, 3, 1),       # rows wider than one 8 KB segment (fp32: 2 segments of 32 pixels)
    (1, 20, 72, 64, 8, 3, 1),       # 3 segments, the last one short
    (2, 9, 136, 20, 20, 5, 0),      # 80 B pixels: segment = 102 pixels rounded to a 16-byte multiple
    (5, 50, 32, 8, 8, 3, 1),        # more strips than ring slots, rows above / below the image in most strips
    (2, 31, 33, 16, 64, 2, 0),      # space-to-depth form of a discriminator layer
]


@pytest.mark.parametrize("case", RAW_CASES)
@pytest.mark.parametrize("xdt", ["f32", "bf16"])
def test_conv_nc_raw_staging_matches_register_staging(ops, case, xdt, monkeypatch):
    """rows brought in by cp.async.bulk + converter warps (DAFK_NC_RAW=1, the default where the layout allows it) must
    give the register-staged kernel's bits: same bf16 rounding, same raster, same MMA order.  The weight gradient ends in
    atomics across CTAs, so it is compared at the accumulation-order tolerance and against the oracle."""
    from multimodal_segmentation_b200._lib import ACT_LRELU
    N, H, W, Cin, Cout, k, pad = case
    r, x, w, b = _mk(case, sum(case) + 11)
    dt = torch.float32 if xdt == "f32" else torch.bfloat16
    wp = ops.pack_conv_nc(gpu(w), 0)
    xg = gpu(x, dt)
    Ho, Wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    dy = bf16_round(r.normal(size=(N, Ho, Wo, Cout)).astype(np.float32))
    dyg = gpu(dy, dt)
    wt = torch.zeros(k, k, Cin, Cout, dtype=torch.float64, requires_grad=True)
    (R.conv2d(t(x, torch.float64), wt, None, 1, "same" if pad else "valid") * t(dy, torch.float64)).sum().backward()
    out = {}
    for raw in ("0", "1", "1"):             # the repeat: nothing may depend on leftovers of an earlier launch
        monkeypatch.setenv("DAFK_NC_RAW", raw)
        y = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad)
        a = ops.conv_nc_fwd(xg, wp, gpu(b), Cout, k, k, pad, ACT_LRELU, 0.3, torch.bfloat16)
        dw, db = ops.zeros(k, k, Cin, Cout), ops.zeros(Cout)
        ops.conv_nc_wgrad(xg, dyg, dw, db, pad)
        torch.cuda.synchronize()
        if raw == "0":
            out = dict(y=y, a=a, dw=dw, db=db)
            continue
        assert torch.equal(out["y"], y) and torch.equal(out["a"], a)
        assert rel_l2(cpu(dw), cpu(out["dw"])) < 1e-5 and rel_l2(cpu(db), cpu(out["db"])) < 1e-5
        assert rel_l2(cpu(dw), wt.grad.numpy()) < 1e-3
        assert rel_l2(cpu(db), dy.sum((0, 1, 2))) < 1e-3
    assert rel_l2(cpu(y), _ref_conv(x, w, b, pad).numpy()) < 1e-4


# ------------------------------------------------------------------ stride-2 valid layers through space-to-depth
S2_CASES = [
    # N, H, W, Cin, Cout, k     (models/discriminator.py:24 first layer; model_components/modality_encoder.py:36-42)
    (2, 32, 32, 1, 64, 4),
    (2, 32, 40, 4, 64, 4),
    (2, 32, 32, 9, 16, 3),
    (2, 31, 31, 16, 32, 3),      # odd input size: zero-padded 2x2 blocks
    (3, 27, 27, 32, 64, 3),      # 4*Cin = 128: tcgen05 kernels on the space-to-depth tensor
    (2, 13, 13, 64, 128, 3),
]


@pytest.mark.parametrize("case", S2_CASES)
def test_stride2_conv_through_space_to_depth(ops, case):
    from multimodal_segmentation_b200 import engine as E
    N, H, W, Cin, Cout, k = case
    r = np.random.RandomState(sum(case))
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    w = bf16_round((r.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32))
    b = r.normal(size=Cout).astype(np.float32)
    xt, wt = t(x, torch.float64, grad=True), t(w, torch.float64, grad=True)
    yr = R.leaky_relu(R.conv2d(xt, wt, t(b, torch.float64), 2, "valid"), 0.3)
    dy = bf16_round(r.normal(size=tuple(yr.shape)).astype(np.float32))
    (yr * t(dy, torch.float64)).sum().backward()
    E.USE_TC = True
    arena = E.Arena(True)
    conv = E.Conv2D(arena, r, "c", Cin, Cout, k, 2, "valid")
    arena.to_device()
    conv.kernel.data.copy_(gpu(w))
    conv.bias.data.copy_(gpu(b))
    tape = E.Tape()
    ctx = E.Ctx(tape, True)
    xv = E.Var(gpu(x), True)
    # narrow layers go to the raster-strip kernels, 4*Cin % 64 == 0 and Cout % 64 == 0 to the swizzled tcgen05 kernels
    assert conv.s2d_eligible([xv])
    y = conv(ctx, xv, "lrelu", 0.3)
    assert tuple(y.shape) == tuple(yr.shape)
    assert rel_l2(cpu(y.data), yr.detach().numpy()) < 1e-4
    y.grad = gpu(dy)
    tape.backward()
    # the product rounds dy * act'(y) to bf16 while it stages it: 5e-3 covers that rounding
    assert rel_l2(cpu(xv.grad), xt.grad.numpy()) < 5e-3
    assert rel_l2(cpu(conv.kernel.grad), wt.grad.numpy()) < 5e-3
    assert rel_l2(cpu(conv.bias.grad), (dy * (yr.detach().numpy() > 0) + 0.3 * dy * (yr.detach().numpy() < 0)).sum((0, 1, 2))) < 5e-3


def _s2d_ref(x):
    """y[n, i, j, (dy*2+dx)*C + c] = x[n, 2i+dy, 2j+dx, c], zero beyond odd sizes (csrc/s2d.cu)"""
    N, H, W, C = x.shape
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    xp = np.zeros((N, 2 * H2, 2 * W2, C), x.dtype)
    xp[:, :H, :W] = x
    return xp.reshape(N, H2, 2, W2, 2, C).transpose(0, 1, 3, 2, 4, 5).reshape(N, H2, W2, 4 * C)


@pytest.mark.parametrize("shape", [(2, 9, 11, 9), (3, 14, 14, 4), (2, 13, 12, 16), (2, 7, 8, 64), (1, 5, 5, 1)])
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_space_to_depth_kernels_bit_exact(ops, shape, dt, monkeypatch):
    """both forward kernels (one thread per pixel quadrant = default, one per element) and the inverse, odd and even
    extents, vector (C % 8 == 0) and scalar channel counts"""
    r = np.random.RandomState(sum(shape))
    x = bf16_round(r.normal(size=shape).astype(np.float32))
    xg = gpu(x, torch.bfloat16 if dt == "bf16" else torch.float32)
    want = _s2d_ref(x)
    y = ops.space_to_depth2(xg)
    assert y.dtype == torch.bfloat16 and np.array_equal(cpu(y), want)
    back = ops.depth_to_space2(y, shape[1], shape[2])
    assert np.array_equal(cpu(back), x)


@pytest.mark.parametrize("shape", [(2, 9, 11, 8, 1), (2, 14, 14, 1, 8), (1, 7, 6, 3, 5), (2, 12, 13, 8, 8)])
def test_space_to_depth_of_a_concatenation_and_its_split_backward(ops, shape):
    """dafk_space_to_depth2_cat == space_to_depth2(concat) bit for bit (mixed source dtypes); depth_to_space2_split hands
    each source its slice of the rearranged gradient"""
    N, H, W, Ca, Cb = shape
    r = np.random.RandomState(sum(shape))
    a = bf16_round(r.normal(size=(N, H, W, Ca)).astype(np.float32))
    b = bf16_round(r.normal(size=(N, H, W, Cb)).astype(np.float32))
    want = _s2d_ref(np.concatenate([a, b], -1))
    for da, db in ((torch.float32, torch.float32), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32)):
        y = ops.space_to_depth2_cat(gpu(a, da), gpu(b, db))
        assert np.array_equal(cpu(y), want)
    g = bf16_round(r.normal(size=want.shape).astype(np.float32))
    for gdt in (torch.float32, torch.bfloat16):
        ga, gb = ops.depth_to_space2_split(gpu(g, gdt), H, W, Ca, Cb)
        full = cpu(ops.depth_to_space2(gpu(g, gdt), H, W))
        assert np.array_equal(cpu(ga), full[..., :Ca]) and np.array_equal(cpu(gb), full[..., Ca:])
    ga, gb = ops.depth_to_space2_split(gpu(g), H, W, Ca, Cb, want_a=False)
    assert ga is None and np.array_equal(cpu(gb), full[..., Ca:])


def test_modality_encoder_first_layer_reads_two_sources(ops):
    """Concatenate([anatomy, image]) -> Conv2D(3x3, stride 2) (model_components/modality_encoder.py:34-38): the fused
    two-source path against the oracle convolution of the concatenation, forward and all gradients"""
    from multimodal_segmentation_b200 import engine as E
    N, H, W, Ca, Cb, Cout, k = 2, 33, 31, 8, 1, 16, 3
    r = np.random.RandomState(5)
    a = bf16_round(r.normal(size=(N, H, W, Ca)).astype(np.float32))
    b = bf16_round(r.normal(size=(N, H, W, Cb)).astype(np.float32))
    w = bf16_round((r.normal(size=(k, k, Ca + Cb, Cout)) / np.sqrt(k * k * (Ca + Cb))).astype(np.float32))
    bias = r.normal(size=Cout).astype(np.float32)
    at, bt, wt = t(a, torch.float64, grad=True), t(b, torch.float64, grad=True), t(w, torch.float64, grad=True)
    yr = R.leaky_relu(R.conv2d(torch.cat([at, bt], -1), wt, t(bias, torch.float64), 2, "valid"), 0.3)
    dy = bf16_round(r.normal(size=tuple(yr.shape)).astype(np.float32))
    (yr * t(dy, torch.float64)).sum().backward()
    E.USE_TC = True
    arena = E.Arena(True)
    conv = E.Conv2D(arena, r, "c", Ca + Cb, Cout, k, 2, "valid")
    arena.to_device()
    conv.kernel.data.copy_(gpu(w))
    conv.bias.data.copy_(gpu(bias))
    tape = E.Tape()
    ctx = E.Ctx(tape, True)
    av, bv = E.Var(gpu(a), True), E.Var(gpu(b), True)
    n0 = ops._lib.launch_count()
    y = conv(ctx, [av, bv], "lrelu", 0.3)
    assert ops._lib.launch_count() - n0 <= 5           # rearrangement + weight operands + convolution, no concatenation
    assert rel_l2(cpu(y.data), yr.detach().numpy()) < 1e-4
    y.grad = gpu(dy)
    tape.backward()
    assert rel_l2(cpu(av.grad), at.grad.numpy()) < 5e-3 and rel_l2(cpu(bv.grad), bt.grad.numpy()) < 5e-3
    assert rel_l2(cpu(conv.kernel.grad), wt.grad.numpy()) < 5e-3


@pytest.mark.parametrize("case", [(2, 20, 24, 8, 8, 20, 5, 0, "f32"), (3, 17, 19, 8, 8, 20, 5, 0, "bf16"), (2, 12, 14, 16, 8, 16, 3, 1, "f32"),
                                  (2, 12, 14, 8, 4, 8, 3, 1, "f32")])
def test_conv_nc_two_concatenated_sources(ops, case):
    """dafk_conv_nc_fwd_cat / dafk_conv_nc_wgrad_cat (the locnet's Concatenate([s1, s2]) -> Conv2D(20, 5),
    layers/stn_spline.py:104-106): bit-identical to the single-source kernels on the materialised concatenation"""
    from multimodal_segmentation_b200._lib import ACT_LRELU
    N, H, W, Ca, Cb, Cout, k, pad, dt = case
    tdt = torch.bfloat16 if dt == "bf16" else torch.float32
    r = np.random.RandomState(sum(case[:8]))
    a = gpu(bf16_round(r.normal(size=(N, H, W, Ca)).astype(np.float32)), tdt)
    b = gpu(bf16_round(r.normal(size=(N, H, W, Cb)).astype(np.float32)), tdt)
    w = gpu((r.normal(size=(k, k, Ca + Cb, Cout)) / np.sqrt(k * k * (Ca + Cb))).astype(np.float32))
    bias = gpu(r.normal(size=Cout).astype(np.float32))
    cat = torch.cat([a, b], -1).contiguous()
    assert ops.nc_supported(Ca + Cb, Cout, k, k, W, pad, 0) and ops.nc_supported(Ca + Cb, Cout, k, k, W, pad, 2)
    wp = ops.pack_conv_nc(w, 0)
    y1 = ops.conv_nc_fwd(cat, wp, bias, Cout, k, k, pad, ACT_LRELU, 0.3)
    y2 = ops.conv_nc_fwd_cat(a, b, wp, bias, Cout, k, k, pad, ACT_LRELU, 0.3)
    assert torch.equal(y1, y2)
    dy = gpu(r.normal(size=tuple(y1.shape)).astype(np.float32))
    dw1, db1, dw2, db2 = ops.zeros(k, k, Ca + Cb, Cout), ops.zeros(Cout), ops.zeros(k, k, Ca + Cb, Cout), ops.zeros(Cout)
    ops.conv_nc_wgrad(cat, dy, dw1, db1, pad)
    ops.conv_nc_wgrad_cat(a, b, dy, dw2, db2, pad)
    torch.cuda.synchronize()
    # the weight gradient accumulates per-CTA partial sums with atomics: same terms, order may differ
    assert rel_l2(cpu(dw2), cpu(dw1)) < 1e-5 and rel_l2(cpu(db2), cpu(db1)) < 1e-5


@pytest.mark.parametrize("case", [
    (8, 220, 220, 20, 16, 5, 4),        # locnet data gradient: 3 segments per row, 3 channel groups, one raster stage
    (8, 224, 224, 64, 8, 3, 1),         # 64 -> 8 data gradient: 7 segments per row, 8 channel groups
    (8, 224, 224, 8, 8, 3, 1),          # FiLM layer: one segment per row
])
def test_conv_nc_raw_staging_repeated_launches_are_identical(ops, case, monkeypatch):
    """200 launches with bulk-copy staging forced must all give the first launch's bits.  Bulk copies complete out of order:
    with a ring whose slots were shared between converter warps a warp could pass a slot's parity wait on a stale phase and
    convert a segment that had not arrived (found by scripts/stress_nc.py as a rare trapped launch); the ring now has a
    multiple of the converter warps as slot count, so that every slot has one owner."""
    N, H, W, Cin, Cout, k, pad = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(k, k, Cin, Cout, device="cuda", generator=g) * 0.1
    wp = ops.pack_conv_nc(w, 0)
    monkeypatch.setenv("DAFK_NC_RAW", "0")
    ref = ops.conv_nc_fwd(x, wp, None, Cout, k, k, pad)
    monkeypatch.setenv("DAFK_NC_RAW", "1")
    outs = [ops.conv_nc_fwd(x, wp, None, Cout, k, k, pad) for _ in range(200)]
    torch.cuda.synchronize()
    bad = [i for i, o in enumerate(outs) if not torch.equal(o, ref)]
    assert not bad, bad[:10]

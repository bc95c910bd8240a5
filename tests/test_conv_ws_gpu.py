"""Parity of the warp-strip narrow-channel convolutions (csrc/conv_ws.cu: bf16 raster in shared memory, mma.sync,
accumulators in registers) against the oracle.

Operands are rounded to bf16 while they are staged, accumulation is fp32; the oracle is evaluated in fp64 on the SAME
bf16-rounded operands, so the tolerance (1e-4 forward / data gradient, 1e-3 for the atomically reduced weight
gradient) only covers the accumulation order.  Against the unrounded fp32 oracle the bound is the north-star 1e-2.
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
    (3, 37, 45, 8, 8, 3, 1),        # ragged strips / chunks
    (2, 24, 40, 8, 64, 3, 1),       # segmentor conv1 (model_components/segmentor.py:15)
    (2, 24, 24, 1, 64, 3, 1),       # UNet first layer (models/unet.py:95)
    (2, 24, 40, 64, 8, 3, 1),       # data gradient of segmentor conv1 as a forward problem (8 channel groups)
    (2, 40, 36, 16, 20, 5, 0),      # locnet conv1 (layers/stn_spline.py:106)
    (2, 30, 30, 20, 20, 5, 0),      # locnet conv2/3
    (2, 20, 28, 8, 1, 1, 0),        # decoder output 1x1 (decoder.py:28)
    (1, 224, 224, 8, 8, 3, 1),      # full-resolution strip geometry
    (2, 12, 12, 9, 16, 3, 1),       # odd channel count (scalar staging path)
    (2, 18, 22, 4, 5, 3, 1),
    (2, 31, 33, 4, 64, 2, 0),       # space-to-depth form of a 4x4 stride-2 first layer (models/discriminator.py:24)
    (2, 31, 33, 16, 64, 2, 0),
    (2, 28, 28, 36, 16, 2, 0),      # modality encoder, space-to-depth (model_components/modality_encoder.py:36-42)
    (2, 14, 15, 64, 32, 2, 0),
    (2, 13, 13, 32, 64, 2, 1),
    (70, 40, 24, 8, 8, 3, 1),       # many strips per CTA: the raster is reused, rows outside the image re-zeroed
    (2, 20, 20, 8, 48, 3, 1),       # 6 n-tiles run on the 8-tile instantiation
]


def _mk(case, seed):
    N, H, W, Cin, Cout, k, pad = case
    r = np.random.RandomState(seed)
    x = bf16_round(r.normal(size=(N, H, W, Cin)).astype(np.float32))
    w = bf16_round((r.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)).astype(np.float32))
    b = r.normal(size=Cout).astype(np.float32)
    return r, x, w, b


def _pad_mode(k, pad):
    return "same" if pad == k // 2 and k % 2 == 1 and pad > 0 else "valid"


def _ref_conv(x, w, b, k, pad):
    xt = x if torch.is_tensor(x) else t(x, torch.float64)
    if pad and _pad_mode(k, pad) == "valid":          # general zero padding
        xt = torch.nn.functional.pad(xt, (0, 0, pad, pad, pad, pad))
    return R.conv2d(xt, w if torch.is_tensor(w) else t(w, torch.float64), None if b is None else t(b, torch.float64), 1,
                    _pad_mode(k, pad))


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("xdt", ["f32", "bf16"])
def test_conv_ws_forward(ops, case, xdt):
    from multimodal_segmentation_b200._lib import ACT_LRELU, ACT_RELU, ACT_TANH
    N, H, W, Cin, Cout, k, pad = case
    assert ops.ws_supported(Cin, Cout, k, k, W, pad, 0)
    r, x, w, b = _mk(case, sum(case))
    yr = _ref_conv(x, w, b, k, pad).numpy()
    xg = gpu(x, torch.float32 if xdt == "f32" else torch.bfloat16)
    for _ in range(2):                                # second launch: nothing may depend on leftovers of the first
        y = ops.conv_ws_fwd(xg, gpu(w), gpu(b), pad)
        assert tuple(y.shape) == tuple(yr.shape)
        assert rel_l2(cpu(y), yr) < 1e-4
    # fused epilogues, bf16 output
    ya = ops.conv_ws_fwd(xg, gpu(w), gpu(b), pad, ACT_LRELU, 0.3, torch.bfloat16)
    assert ya.dtype == torch.bfloat16 and rel_l2(cpu(ya), np.where(yr > 0, yr, 0.3 * yr)) < 5e-3
    yt = ops.conv_ws_fwd(xg, gpu(w), None, pad, ACT_TANH)
    assert rel_l2(cpu(yt), np.tanh(_ref_conv(x, w, None, k, pad).numpy())) < 1e-4
    # folded inference BatchNorm: per-output-channel scale on the weights (rounded to bf16 AFTER scaling) + ReLU
    sc = r.uniform(0.5, 2.0, size=Cout).astype(np.float32)
    ws = bf16_round(w * sc)
    yf = ops.conv_ws_fwd(xg, gpu(w), gpu(b), pad, ACT_RELU, 0.0, torch.float32, scale=gpu(sc))
    assert rel_l2(cpu(yf), np.maximum(_ref_conv(x, ws, b, k, pad).numpy(), 0)) < 1e-4


@pytest.mark.parametrize("case", CASES)
def test_conv_ws_dgrad(ops, case):
    N, H, W, Cin, Cout, k, pad = case
    if not ops.ws_supported(Cout, Cin, k, k, W + 2 * pad - k + 1, k - 1 - pad, 1):
        pytest.skip("data gradient of this layer is not a warp-strip shape")
    r, x, w, b = _mk(case, sum(case) + 1)
    xt = torch.zeros(N, H, W, Cin, dtype=torch.float64, requires_grad=True)
    yr = _ref_conv(xt, t(w, torch.float64), None, k, pad)
    dy = bf16_round(r.normal(size=tuple(yr.shape)).astype(np.float32))
    (yr * t(dy, torch.float64)).sum().backward()
    dx = ops.conv_ws_fwd(gpu(dy), gpu(w), None, k - 1 - pad, mode=1)
    assert tuple(dx.shape) == (N, H, W, Cin)
    assert rel_l2(cpu(dx), xt.grad.numpy()) < 1e-4
    # activation backward fused into the staging: dy * lrelu'(y) with y the layer's activation output
    from multimodal_segmentation_b200._lib import ACT_LRELU
    yact = r.normal(size=tuple(yr.shape)).astype(np.float32)
    yact[0, 0, 0, 0] = 0.0                                              # derivative at exactly 0 is 0 (Keras 2.1.6)
    g = bf16_round(dy * np.where(yact > 0, 1.0, np.where(yact < 0, 0.3, 0.0)).astype(np.float32))
    xt2 = torch.zeros(N, H, W, Cin, dtype=torch.float64, requires_grad=True)
    (_ref_conv(xt2, t(w, torch.float64), None, k, pad) * t(g, torch.float64)).sum().backward()
    for adt in (torch.float32, torch.bfloat16):
        ya = gpu(yact, adt)
        if adt == torch.bfloat16:                                       # signs survive the rounding
            assert np.array_equal(np.sign(cpu(ya)), np.sign(yact))
        dx2 = ops.conv_ws_fwd(gpu(dy), gpu(w), None, k - 1 - pad, mode=1, ya=ya, gact=ACT_LRELU, galpha=0.3)
        assert rel_l2(cpu(dx2), xt2.grad.numpy()) < 1e-4


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dts", [("f32", "f32"), ("bf16", "bf16"), ("bf16", "f32")])
def test_conv_ws_wgrad(ops, case, dts):
    N, H, W, Cin, Cout, k, pad = case
    assert ops.ws_supported(Cin, Cout, k, k, W, pad, 2)
    r, x, w, b = _mk(case, sum(case) + 2)
    wt = torch.zeros(k, k, Cin, Cout, dtype=torch.float64, requires_grad=True)
    yr = _ref_conv(x, wt, None, k, pad)
    dy = bf16_round(r.normal(size=tuple(yr.shape)).astype(np.float32))
    (yr * t(dy, torch.float64)).sum().backward()
    dw = ops.zeros(k, k, Cin, Cout)
    db = ops.zeros(Cout)
    td = {"f32": torch.float32, "bf16": torch.bfloat16}
    ops.conv_ws_wgrad(gpu(x, td[dts[0]]), gpu(dy, td[dts[1]]), dw, db, pad)
    assert rel_l2(cpu(dw), wt.grad.numpy()) < 1e-3
    assert rel_l2(cpu(db), dy.sum((0, 1, 2))) < 1e-3
    # accumulates into dw (second call doubles it), db optional
    ops.conv_ws_wgrad(gpu(x, td[dts[0]]), gpu(dy, td[dts[1]]), dw, None, pad)
    assert rel_l2(cpu(dw), 2 * wt.grad.numpy()) < 1e-3
    assert rel_l2(cpu(db), dy.sum((0, 1, 2))) < 1e-3


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[5], CASES[13]])
def test_conv_ws_wgrad_with_fused_activation_backward(ops, case):
    from multimodal_segmentation_b200._lib import ACT_LRELU, ACT_RELU
    N, H, W, Cin, Cout, k, pad = case
    r, x, w, b = _mk(case, sum(case) + 4)
    wt = torch.zeros(k, k, Cin, Cout, dtype=torch.float64, requires_grad=True)
    yr = _ref_conv(x, wt, None, k, pad)
    dy = bf16_round(r.normal(size=tuple(yr.shape)).astype(np.float32))
    yact = r.normal(size=tuple(yr.shape)).astype(np.float32)
    for act, alpha in ((ACT_LRELU, 0.2), (ACT_RELU, 0.0)):
        g = bf16_round(dy * np.where(yact > 0, 1.0, alpha).astype(np.float32))
        wt.grad = None
        (yr * t(g, torch.float64)).sum().backward(retain_graph=True)
        dw, db = ops.zeros(k, k, Cin, Cout), ops.zeros(Cout)
        ops.conv_ws_wgrad(gpu(x), gpu(dy), dw, db, pad, ya=gpu(yact), gact=act, galpha=alpha)
        assert rel_l2(cpu(dw), wt.grad.numpy()) < 1e-3
        assert rel_l2(cpu(db), g.sum((0, 1, 2))) < 1e-3


def test_conv_ws_vs_fp32_oracle_tolerance(ops):
    """north-star bound for the bf16 conv path against the unrounded fp32 reference: <= 1e-2"""
    r = np.random.RandomState(3)
    x = r.normal(size=(2, 32, 32, 8)).astype(np.float32)
    w = (r.normal(size=(3, 3, 8, 8)) / np.sqrt(72)).astype(np.float32)
    yr = _ref_conv(x, w, None, 3, 1).numpy()
    y = ops.conv_ws_fwd(gpu(x), gpu(w), None, 1)
    assert rel_l2(cpu(y), yr) < 1e-2


def test_conv_ws_matches_raster_strip_kernel_bits(ops):
    """same bf16 operands, fp32 accumulation: the warp-strip and the tcgen05 raster-strip kernels agree to fp32 round-off"""
    r, x, w, b = _mk((2, 24, 40, 8, 8, 3, 1), 5)
    y0 = ops.conv_nc_fwd(gpu(x), ops.pack_conv_nc(gpu(w), 0), gpu(b), 8, 3, 3, 1)
    y1 = ops.conv_ws_fwd(gpu(x), gpu(w), gpu(b), 1)
    assert rel_l2(cpu(y1), cpu(y0)) < 1e-6

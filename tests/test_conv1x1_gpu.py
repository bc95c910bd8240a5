"""Parity of the pointwise 64 -> <=8 head kernels (csrc/conv_1x1.cu: anatomy head 64 -> 8,
model_components/anatomy_encoder.py:26; segmentor head 64 -> 5, model_components/segmentor.py:24) against the oracle.

The kernels multiply bf16 operands (the feature map is stored in bf16; weights and the incoming gradient are rounded to
bf16 as they are loaded) and accumulate in fp32; the oracle runs in fp64 on the SAME rounded operands, so the tolerances
(1e-5 forward, 1e-4 for the bf16-stored data gradient = its storage rounding, 1e-4 for the atomically reduced weight
gradient) only cover accumulation order and output storage.  Sizes cover whole 256-pixel tiles (bulk-copy pipeline),
ragged tails, fewer pixels than one tile and the empty input.
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
    # N, H, W, Cout
    (2, 32, 32, 8),        # 8 whole tiles
    (2, 32, 32, 5),
    (3, 37, 45, 5),        # ragged tail, Cout*4 B rows not 16 B aligned
    (3, 37, 45, 8),
    (1, 9, 11, 5),         # fewer pixels than one tile
    (1, 9, 11, 1),
    (4, 224, 224, 8),      # more tiles than CTAs: the ring wraps
    (4, 224, 224, 5),
]


def _mk(case, seed):
    N, H, W, Cout = case
    r = np.random.RandomState(seed)
    x = bf16_round(r.normal(size=(N, H, W, 64)).astype(np.float32))
    w = (r.normal(size=(1, 1, 64, Cout)) / 8.0).astype(np.float32)
    b = r.normal(size=Cout).astype(np.float32)
    dy = r.normal(size=(N, H, W, Cout)).astype(np.float32)
    return x, w, b, dy


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("rnd", [True, False])
def test_conv1x1_forward(ops, case, rnd):
    x, w, b, _ = _mk(case, 1)
    wr = bf16_round(w) if rnd else w
    y = ops.conv1x1_fwd(gpu(x, torch.bfloat16), gpu(w), gpu(b), round_bf16=rnd)
    ref = R.conv2d(t(x, torch.float64), t(wr, torch.float64), t(b, torch.float64), 1, "valid").numpy()
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert rel_l2(cpu(y), ref) < 1e-5
    # against the unrounded fp32 oracle: the north-star bf16 bound
    ref32 = R.conv2d(t(x), t(w), t(b), 1, "valid").numpy()
    assert rel_l2(cpu(y), ref32) < 1e-2


@pytest.mark.parametrize("case", CASES)
def test_conv1x1_backward(ops, case):
    x, w, b, dy = _mk(case, 2)
    N, H, W, Cout = case
    xt = t(x, torch.float64, grad=True)
    wt = t(bf16_round(w), torch.float64, grad=True)
    bt = t(b, torch.float64, grad=True)
    y = R.conv2d(xt, wt, bt, 1, "valid")
    # the kernels consume the output gradient rounded to bf16 (bias gradient: unrounded column sum)
    y.backward(t(bf16_round(dy), torch.float64))
    dx = ops.conv1x1_dgrad(gpu(dy), gpu(w))
    assert dx.dtype == torch.bfloat16 and tuple(dx.shape) == (N, H, W, 64)
    assert rel_l2(cpu(dx), xt.grad.numpy()) < 4e-3          # bf16 storage of dx (2^-9 per element)
    assert rel_l2(cpu(dx), bf16_round(xt.grad.numpy().astype(np.float32))) < 1e-3   # up to 1-ulp flips
    dw = torch.zeros(1, 1, 64, Cout, device="cuda")
    db = torch.zeros(Cout, device="cuda")
    ops.conv1x1_wgrad(gpu(x, torch.bfloat16), gpu(dy), dw, db)
    assert rel_l2(cpu(dw), wt.grad.numpy()) < 1e-4
    assert rel_l2(cpu(db), dy.reshape(-1, Cout).astype(np.float64).sum(0)) < 1e-5
    # accumulation semantics: a second call doubles the gradient
    ops.conv1x1_wgrad(gpu(x, torch.bfloat16), gpu(dy), dw, None)
    assert rel_l2(cpu(dw), 2.0 * wt.grad.numpy()) < 1e-4


def test_conv1x1_empty_and_errors(ops):
    from multimodal_segmentation_b200._lib import DafkError
    x = torch.empty(0, 4, 4, 64, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(1, 1, 64, 5, device="cuda")
    y = ops.conv1x1_fwd(x, w, None)
    assert tuple(y.shape) == (0, 4, 4, 5)
    assert not ops.conv1x1_supported(32, 5) and not ops.conv1x1_supported(64, 9)
    with pytest.raises(DafkError):
        ops.conv1x1_fwd(torch.zeros(1, 2, 2, 32, dtype=torch.bfloat16, device="cuda"),
                        torch.zeros(1, 1, 32, 5, device="cuda"), None)

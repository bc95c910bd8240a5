"""Thin, allocation-only wrappers over the C-ABI kernels (``include/dafk.h``).

Every function here takes CUDA torch tensors (used purely as device-memory handles), allocates
the outputs with the caching allocator and enqueues ONE OR MORE hand-written kernels on torch's
current stream through ctypes.  No arithmetic is done by torch.  There is no autograd here --
the reverse pass is recorded by ``tape.py``.
"""
import os

import torch

from . import _lib, instrument
from ._lib import ACT_LRELU, ACT_NONE, DAFK_BF16, DAFK_F32, ConvDesc, call

_S = _lib.stream_ptr


def _dt(t):
    if t.dtype == torch.float32:
        return DAFK_F32
    if t.dtype == torch.bfloat16:
        return DAFK_BF16
    raise TypeError("unsupported dtype %s" % t.dtype)


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.DafkError("dafk kernels need CUDA tensors; there is no CPU fallback")
        if not t.is_contiguous():
            raise _lib.DafkError("dafk kernels need contiguous tensors")


def f32(*shape, device="cuda"):
    return torch.empty(*shape, dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------- pointwise
def round_fwd(x):
    _chk(x)
    y = torch.empty_like(x)
    call("round_fwd", x, y, x.numel(), _S())
    return y


def softmax_fwd(x, want_round=False):
    _chk(x)
    C = x.shape[-1]
    p = torch.empty_like(x)
    r = torch.empty_like(x) if want_round else None
    call("softmax_fwd", x, p, r, x.numel() // C, C, _S())
    return p, r


def softmax_bwd(p, dp):
    _chk(p, dp)
    C = p.shape[-1]
    dx = torch.empty_like(p)
    call("softmax_bwd", p, dp, dx, p.numel() // C, C, _S())
    return dx


def act_fwd(x, act, alpha=0.0, inplace=False):
    _chk(x)
    y = x if inplace else torch.empty_like(x)
    call("act_fwd", x, y, x.numel(), act, float(alpha), _S())
    return y


def act_bwd(dy, y, act, alpha=0.0, inplace=False):
    _chk(dy, y)
    dx = dy if inplace else torch.empty_like(dy)
    call("act_bwd", dy, y, dx, y.numel(), act, float(alpha), _S())
    return dx


def add_act_bwd(dy1, dy2, y, act, alpha=0.0, out=None):
    """(dy1 + dy2) * act'(y) in one pass (fp32); ``out`` may be dy1 or dy2"""
    _chk(dy1, dy2, y)
    assert dy1.dtype == dy2.dtype == y.dtype == torch.float32 and dy1.shape == dy2.shape == y.shape
    dx = torch.empty_like(dy1) if out is None else out
    call("add_act_bwd", dy1, dy2, y, dx, y.numel(), act, float(alpha), _S())
    return dx


def act_bwd_bf16(dy, y, act, alpha=0.0):
    """bf16(act'(y) * dy) in one pass (fp32 in)"""
    _chk(dy, y)
    dx = torch.empty(dy.shape, dtype=torch.bfloat16, device=dy.device)
    call("act_bwd_bf16", dy, y, dx, y.numel(), act, float(alpha), _S())
    return dx


def act_bwd_bf16io(dy, y, act, alpha=0.0):
    """act'(y) * dy with bf16 gradient and bf16 activation output"""
    _chk(dy, y)
    assert dy.dtype == y.dtype == torch.bfloat16
    dx = torch.empty_like(dy)
    call("act_bwd_bf16io", dy, y, dx, y.numel(), act, float(alpha), _S())
    return dx


def add(a, b, out=None):
    _chk(a, b)
    out = torch.empty_like(a) if out is None else out
    call("add_dt", a, b, out, _dt(a), a.numel(), _S())
    return out


def add_(a, b):
    """a += b"""
    return add(a, b, out=a)


def axpby_(a, x, b, y):
    """y = a*x + b*y"""
    _chk(x, y)
    call("axpby", float(a), x, float(b), y, x.numel(), _S())
    return y


def fill_(x, v):
    _chk(x)
    call("fill", x, float(v), x.numel(), _S())
    return x


def zero_(t):
    """cudaMemsetAsync on the current stream (no torch kernel)."""
    _chk(t)
    call("memset_zero", t, t.numel() * t.element_size(), _S())
    return t


def zeros(*shape, dtype=torch.float32):
    return zero_(torch.empty(*shape, dtype=dtype, device="cuda"))


def cast(x, dtype):
    _chk(x)
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    call("cast", x, _dt(x), y, _dt(y), x.numel(), _S())
    return y


def copy_(dst, src):
    """dst <- src (same shape; dtype conversion allowed), our own copy kernel"""
    _chk(dst, src)
    assert dst.numel() == src.numel()
    call("cast", src, _dt(src), dst, _dt(dst), src.numel(), _S())
    return dst


def copy_channels(src, src_off, dst, dst_off, c, accumulate=False):
    _chk(src, dst)
    M = src.numel() // src.shape[-1]
    call("copy_channels", src, src.shape[-1], src_off, dst, dst.shape[-1], dst_off, c, M, int(accumulate), _S())
    return dst


def concat_channels(tensors):
    C = sum(t.shape[-1] for t in tensors)
    out = f32(*tensors[0].shape[:-1], C)
    off = 0
    for t in tensors:
        copy_channels(t, 0, out, off, t.shape[-1])
        off += t.shape[-1]
    return out


def slice_channels(x, off, c):
    out = f32(*x.shape[:-1], c)
    copy_channels(x, off, out, 0, c)
    return out


def gather_rows(src, idx):
    _chk(src, idx)
    rows = idx.numel()
    row_elems = src.numel() // src.shape[0]
    out = torch.empty((rows,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    call("gather_rows", src, idx, out, rows, row_elems, _S())
    return out


# ---------------------------------------------------------------------------- FiLM / Maximum
def film_fwd(x, gamma, beta):
    _chk(x, gamma, beta)
    B, C = x.shape[0], x.shape[-1]
    y = torch.empty_like(x)
    call("film_fwd", x, gamma, beta, y, B, x.numel() // (B * C), C, _S())
    return y


def film_bwd(dy, x, gamma):
    _chk(dy, x, gamma)
    B, C = x.shape[0], x.shape[-1]
    dx = torch.empty_like(x)
    dg = f32(B, C)
    db = f32(B, C)
    ws = torch.empty(B * C * 2, dtype=torch.float64, device=x.device)
    call("film_bwd", dy, x, gamma, dx, dg, db, ws, B, x.numel() // (B * C), C, _S())
    return dx, dg, db


def film_act_add_fwd(x, gamma, beta, res, act, alpha=0.0):
    """res + act(x*gamma + beta) in one pass (decoder.py:50-54); x / res f32, or both bf16 (gamma, beta stay f32)"""
    _chk(x, gamma, beta, res)
    B, C = x.shape[0], x.shape[-1]
    y = torch.empty_like(x)
    name = "film_act_add_fwd_bf16" if x.dtype == torch.bfloat16 else "film_act_add_fwd"
    assert res.dtype == x.dtype and gamma.dtype == beta.dtype == torch.float32
    call(name, x, gamma, beta, res, y, B, x.numel() // (B * C), C, act, float(alpha), _S())
    return y


def film_act_add_bwd(dy, x, gamma, beta, act, alpha=0.0):
    _chk(dy, x, gamma, beta)
    B, C = x.shape[0], x.shape[-1]
    dx = torch.empty_like(x)
    dg = f32(B, C)
    db = f32(B, C)
    ws = torch.empty(B * C * 2, dtype=torch.float64, device=x.device)
    name = "film_act_add_bwd_bf16" if x.dtype == torch.bfloat16 else "film_act_add_bwd"
    assert dy.dtype == x.dtype and gamma.dtype == beta.dtype == torch.float32
    call(name, dy, x, gamma, beta, dx, dg, db, ws, B, x.numel() // (B * C), C, act, float(alpha), _S())
    return dx, dg, db


def max_fwd(a, b):
    _chk(a, b)
    out = torch.empty_like(a)
    call("max_fwd", a, b, out, a.numel(), _S())
    return out


def max_bwd(a, b, dout):
    _chk(a, b, dout)
    da, db = torch.empty_like(a), torch.empty_like(a)
    call("max_bwd", a, b, dout, da, db, a.numel(), _S())
    return da, db


# ---------------------------------------------------------------------------- batch norm
_bn_ws_cache = {}


def _bn_ws(device):
    """persistent self-resetting workspace of the wide BatchNorm reductions (zeroed once; every launch leaves it
    zero), one per device: all kernels of a rank run on one stream"""
    key = str(device)
    if key not in _bn_ws_cache:
        n = int(_lib.lib().fn["dafk_bn_wide_ws_bytes"](1024))
        _bn_ws_cache[key] = zero_(torch.empty(n, dtype=torch.uint8, device=device))
    return _bn_ws_cache[key]


def _bn_wide(C, M):
    return M > 0 and bool(_lib.lib().fn["dafk_bn_wide_supported"](C))


def bn_stats_finalize(x, eps, momentum, moving_mean=None, moving_var=None):
    _chk(x)
    C = x.shape[-1]
    M = x.numel() // C
    if _bn_wide(C, M):
        ws = _bn_ws(x.device)
        mean, rstd = f32(C), f32(C)
        instrument.timed("bn_stats", 0, float(x.element_size()) * x.numel(),
                         lambda: call("bn_stats_fused", x, _dt(x), ws, ws.numel(), M, C, float(eps), float(momentum), mean,
                                      rstd, moving_mean, moving_var, _S()))
        return mean, rstd
    acc = torch.empty(2 * C, dtype=torch.float64, device=x.device)
    zero_(acc)
    instrument.timed("bn_stats", 0, float(x.element_size()) * x.numel(), lambda: call("bn_stats", x, _dt(x), acc, M, C, _S()))
    mean, rstd = f32(C), f32(C)
    call("bn_finalize", acc, M, C, float(eps), float(momentum), mean, rstd, moving_mean, moving_var, _S())
    return mean, rstd


def bn_rstd_from_var(var, eps):
    rstd = torch.empty_like(var)
    call("bn_rstd_from_var", var, rstd, var.numel(), float(eps), _S())
    return rstd


def bn_apply(x, mean, rstd, gamma, beta, act=ACT_NONE, out_dtype=torch.float32):
    _chk(x)
    C = x.shape[-1]
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    instrument.timed("bn_apply", 0, float(x.element_size()) * x.numel() + out.numel() * out.element_size(),
                     lambda: call("bn_apply", x, _dt(x), mean, rstd, gamma, beta, out, _dt(out), x.numel() // C, C, act, _S()))
    return out


def bn_bwd(dout, x, mean, rstd, gamma, beta, act, dgamma, dbeta, dx_dtype=torch.float32, dbias_prev=None):
    """training-mode backward; dgamma/dbeta (and the producing conv's bias gradient) are accumulated into."""
    _chk(dout, x)
    C = x.shape[-1]
    M = x.numel() // C
    acc = torch.empty(2 * C, dtype=torch.float64, device=x.device)
    nb_in = float(x.element_size()) * x.numel() + dout.numel() * dout.element_size()
    if _bn_wide(C, M):
        ws = _bn_ws(x.device)
        instrument.timed("bn_bwd_reduce", 0, nb_in,
                         lambda: call("bn_bwd_reduce_fused", dout, _dt(dout), x, _dt(x), mean, rstd, gamma, beta, acc, ws,
                                      ws.numel(), M, C, act, _S()))
    else:
        zero_(acc)
        instrument.timed("bn_bwd_reduce", 0, nb_in,
                         lambda: call("bn_bwd_reduce", dout, _dt(dout), x, _dt(x), mean, rstd, gamma, beta, acc, M, C, act, _S()))
    dx = torch.empty(x.shape, dtype=dx_dtype, device=x.device)
    instrument.timed("bn_bwd_apply", 0, nb_in + dx.numel() * dx.element_size(),
                     lambda: call("bn_bwd_apply", dout, _dt(dout), x, _dt(x), mean, rstd, gamma, beta, acc, dx, _dt(dx), dgamma,
                                  dbeta, dbias_prev, M, C, act, _S()))
    return dx


def bn_bwd_frozen(dout, x, mean, rstd, gamma, beta, act):
    _chk(dout, x)
    C = x.shape[-1]
    dx = torch.empty_like(x)
    call("bn_bwd_frozen", dout, x, mean, rstd, gamma, beta, dx, x.numel() // C, C, act, _S())
    return dx


# ---------------------------------------------------------------------------- pooling
def maxpool2_fwd(x):
    _chk(x)
    N, H, W, C = x.shape
    y = torch.empty((N, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
    call("maxpool2_fwd", x, y, _dt(x), N, H, W, C, _S())
    return y


def maxpool2_bwd(x, dy):
    _chk(x, dy)
    N, H, W, C = x.shape
    dx = torch.empty_like(x)
    call("maxpool2_bwd", x, dy, dx, _dt(x), N, H, W, C, _S())
    return dx


def upsample2_fwd(x):
    _chk(x)
    N, H, W, C = x.shape
    if C % 4 != 0 and x.dtype == torch.float32:
        return resize_nn_fwd(x, 2 * H, 2 * W)       # any channel count: nearest resize by 2 is the same map
    y = torch.empty((N, 2 * H, 2 * W, C), dtype=x.dtype, device=x.device)
    call("upsample2_fwd", x, y, _dt(x), N, H, W, C, _S())
    return y


def upsample2_bwd(dy):
    _chk(dy)
    N, H2, W2, C = dy.shape
    dx = torch.empty((N, H2 // 2, W2 // 2, C), dtype=dy.dtype, device=dy.device)
    call("upsample2_bwd", dy, dx, _dt(dy), N, H2 // 2, W2 // 2, C, _S())
    return dx


def resize_nn_fwd(x, Ho, Wo):
    _chk(x)
    N, H, W, C = x.shape
    y = f32(N, Ho, Wo, C)
    call("resize_nn_fwd", x, y, N, H, W, C, Ho, Wo, _S())
    return y


def resize_nn_bwd(dy, H, W):
    _chk(dy)
    N, Ho, Wo, C = dy.shape
    dx = f32(N, H, W, C)
    call("resize_nn_bwd", dy, dx, N, H, W, C, Ho, Wo, _S())
    return dx


# ---------------------------------------------------------------------------- convolution
def conv_desc(N, H, W, Cin, Cout, KH, KW, stride, pad):
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    return ConvDesc(N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo)


USE_SMALL = True     # direct kernels for narrow layers (tests switch it off to exercise the general path)


def _small(Cin, Cout, KH, KW):
    return (USE_SMALL and (Cin <= 32 or Cout <= 32)
            and bool(_lib.lib().fn["dafk_conv_small_supported"](Cin, Cout, KH, KW)))


def conv2d_fwd(x, w, bias, stride=1, pad=0, act=ACT_NONE, alpha=0.0):
    _chk(x, w, bias)
    N, H, W, Cin = x.shape
    KH, KW, _, Cout = w.shape
    d = conv_desc(N, H, W, Cin, Cout, KH, KW, stride, pad)
    y = f32(N, d.Ho, d.Wo, Cout)
    fl = 2.0 * N * d.Ho * d.Wo * Cout * KH * KW * Cin
    if _small(Cin, Cout, KH, KW):
        instrument.timed("conv_small_fwd", fl, 4.0 * (x.numel() + y.numel()),
                         lambda: call("conv_small_fwd", d, x, w, bias, y, act, float(alpha), _S()))
        return y
    instrument.timed("conv2d_generic_fwd", fl, 4.0 * (x.numel() + y.numel()),
                     lambda: call("conv2d_fwd", d, x, w, bias, y, act, float(alpha), _S()))
    return y


def conv2d_dgrad(dy, w, x_shape, stride=1, pad=0):
    _chk(dy, w)
    N, H, W, Cin = x_shape
    KH, KW, _, Cout = w.shape
    d = conv_desc(N, H, W, Cin, Cout, KH, KW, stride, pad)
    dx = f32(N, H, W, Cin)
    fl = 2.0 * N * d.Ho * d.Wo * Cout * KH * KW * Cin
    if _small(Cin, Cout, KH, KW) and (stride == 1 or Cin <= 16):
        instrument.timed("conv_small_dgrad", fl, 4.0 * (dy.numel() + dx.numel()),
                         lambda: call("conv_small_dgrad", d, dy, w, dx, _S()))
        return dx
    instrument.timed("conv2d_generic_dgrad", fl, 4.0 * (dy.numel() + dx.numel()),
                     lambda: call("conv2d_dgrad", d, dy, w, dx, _S()))
    return dx


def conv2d_wgrad(x, dy, dw, db, stride=1, pad=0):
    """dw += ..., db += ... (db may be None)"""
    _chk(x, dy, dw, db)
    N, H, W, Cin = x.shape
    KH, KW, _, Cout = dw.shape
    d = conv_desc(N, H, W, Cin, Cout, KH, KW, stride, pad)
    fl = 2.0 * N * d.Ho * d.Wo * Cout * KH * KW * Cin
    if _small(Cin, Cout, KH, KW):
        instrument.timed("conv_small_wgrad", fl, 4.0 * (x.numel() + dy.numel()),
                         lambda: call("conv_small_wgrad", d, x, dy, dw, db, _S()))
        return
    instrument.timed("conv2d_generic_wgrad", fl, 4.0 * (x.numel() + dy.numel()),
                     lambda: call("conv2d_wgrad", d, x, dy, dw, db, _S()))


def colsum_(x, out):
    _chk(x, out)
    C = x.shape[-1]
    call("colsum", x, _dt(x), out, x.numel() // C, C, _S())
    return out


def pack_conv3x3(w_hwio, for_dgrad=False):
    _chk(w_hwio)
    _, _, Cin, Cout = w_hwio.shape
    shape = (9, Cin, Cout) if for_dgrad else (9, Cout, Cin)
    wp = torch.empty(shape, dtype=torch.bfloat16, device=w_hwio.device)
    call("pack_conv3x3", w_hwio, wp, Cin, Cout, int(for_dgrad), _S())
    return wp


def conv3x3_tc_fwd(x0, x1, wp, bias, Cout, out_dtype=torch.float32, row_off=0):
    """tcgen05 path; x0 (and optional x1 = second concat source) are bf16 NHWC.  wp is the packed
    weight matrix [9][rows][K]; the call produces output channels rows[row_off : row_off+Cout]."""
    _chk(x0, x1, wp, bias)
    N, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[-1]
    y = torch.empty((N, H, W, Cout), dtype=out_dtype, device=x0.device)
    fl = 2.0 * N * H * W * Cout * 9 * (C0 + C1)
    nb = 2.0 * N * H * W * (C0 + C1) + y.numel() * y.element_size()
    instrument.timed("conv_tc_fwd+dgrad (tcgen05)", fl, nb,
                     lambda: call("conv3x3_tc_fwd", x0, C0, x1, C1, wp, wp.shape[1], row_off, bias, y, _dt(y), N, H, W,
                                  Cout, _S()), tag=(N, H, W, C0 + C1, Cout, 3, 1, str(out_dtype)[6:]))
    return y


def pack_conv(w_hwio, mode, pa=0, pb=0, out=None):
    """mode 0: [taps][Cout][Cin] forward; 1: mirrored [taps][Cin][Cout] (stride-1 dgrad);
    2: stride-2 dgrad parity class (pa,pb): [(KH/2)(KW/2)][Cin][Cout]"""
    _chk(w_hwio)
    KH, KW, Cin, Cout = w_hwio.shape
    cip, cop = (Cin + 63) // 64 * 64, (Cout + 63) // 64 * 64      # zero-padded to whole 64-channel blocks
    if mode == 0:
        shape = (KH * KW, cop, cip)
    elif mode == 1:
        shape = (KH * KW, cip, cop)
    else:
        shape = ((KH // 2) * (KW // 2), cip, cop)
    wp = torch.empty(shape, dtype=torch.bfloat16, device=w_hwio.device) if out is None else out
    call("pack_conv", w_hwio, wp, KH, KW, Cin, Cout, mode, pa, pb, _S())
    return wp


def bn_fold(gamma, beta, moving_mean, moving_var, conv_bias, eps, out=None):
    """inference-mode BatchNorm as a per-channel affine map folded into the producing convolution:
    -> (scale [C], bias' [C]); ``out`` = (scale, bias') buffers to refresh in place"""
    _chk(gamma, beta, moving_mean, moving_var, conv_bias)
    C = gamma.numel()
    scale, bias2 = out if out is not None else (f32(C, device=gamma.device), f32(C, device=gamma.device))
    call("bn_fold", gamma, beta, moving_mean, moving_var, conv_bias, float(eps), scale, bias2, C, _S())
    return scale, bias2


def pack_conv_scaled(w_hwio, scale, out=None):
    """forward operand [taps][Cout_pad][Cin_pad] of w * scale[Cout]"""
    _chk(w_hwio, scale)
    KH, KW, Cin, Cout = w_hwio.shape
    cip, cop = (Cin + 63) // 64 * 64, (Cout + 63) // 64 * 64
    wp = torch.empty((KH * KW, cop, cip), dtype=torch.bfloat16, device=w_hwio.device) if out is None else out
    call("pack_conv_scaled", w_hwio, scale, wp, KH, KW, Cin, Cout, _S())
    return wp


def conv_tc_fwd(x0, x1, wp, bias, Cout, KH, KW, stride, pad, out_dtype=torch.float32, row_off=0, out=None, act=ACT_NONE,
                alpha=0.0):
    """general tcgen05 convolution.  ``out`` may be a strided NHWC view (e.g. dx[:, pa::2, pb::2, :]); its
    spatial extent defines the logical output size."""
    _chk(x0, x1, wp, bias)
    N, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[-1]
    if out is None:
        Ho = (H + 2 * pad - KH) // stride + 1
        Wo = (W + 2 * pad - KW) // stride + 1
        out = torch.empty((N, Ho, Wo, Cout), dtype=out_dtype, device=x0.device)
    Ho, Wo = out.shape[1], out.shape[2]
    assert out.stride(3) == 1 and out.shape[3] == Cout
    fl = 2.0 * N * Ho * Wo * Cout * KH * KW * (C0 + C1)
    nb = 2.0 * x0.numel() + (0 if x1 is None else 2.0 * x1.numel()) + N * Ho * Wo * Cout * out.element_size()
    instrument.timed("conv_tc_fwd+dgrad (tcgen05)", fl, nb,
                     lambda: call("conv_tc_fwd_act", x0, C0, x1, C1, wp, wp.shape[1], row_off, bias, out, _dt(out), N, H, W,
                                  Cout, KH, KW, stride, pad, Ho, Wo, out.stride(0), out.stride(1), out.stride(2), int(act),
                                  float(alpha), _S()),
                     tag=(N, H, W, C0 + C1, Cout, KH, stride, str(out.dtype)[6:], Ho))
    return out


FUSE_BN_STATS = os.environ.get("DAFK_FUSE_BN_STATS", "1") != "0"     # batch statistics from the convolution's epilogue


def conv_tc_fwd_bn(x0, x1, wp, bias, Cout, KH, KW, stride, pad):
    """tcgen05 convolution that stores y in bf16 and accumulates sum / sum of squares per channel of the stored values
    into a fresh fp64 accumulator [2*Cout] (for bn_finalize): returns (y, acc)"""
    _chk(x0, x1, wp, bias)
    N, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[-1]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    y = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=x0.device)
    acc = zero_(torch.empty(2 * Cout, dtype=torch.float64, device=x0.device))
    fl = 2.0 * N * Ho * Wo * Cout * KH * KW * (C0 + C1)
    nb = 2.0 * x0.numel() + (0 if x1 is None else 2.0 * x1.numel()) + 2.0 * y.numel()
    instrument.timed("conv_tc_fwd+dgrad (tcgen05)", fl, nb,
                     lambda: call("conv_tc_fwd_bn", x0, C0, x1, C1, wp, wp.shape[1], 0, bias, y, acc, N, H, W, Cout, KH, KW,
                                  stride, pad, _S()),
                     tag=(N, H, W, C0 + C1, Cout, KH, stride, "bfloat16+bn", Ho))
    return y, acc


def bn_finalize_acc(acc, M, eps, momentum, moving_mean=None, moving_var=None):
    """mean / rstd (and the moving-statistics update) from per-channel fp64 sums [2*C] over M values per channel"""
    C = acc.numel() // 2
    mean, rstd = f32(C), f32(C)
    call("bn_finalize", acc, M, C, float(eps), float(momentum), mean, rstd, moving_mean, moving_var, _S())
    return mean, rstd


WGRAD_HALO = os.environ.get("DAFK_WGRAD_HALO", "auto")     # "0" / "1" force the choice (tests, benchmarks)


def pack_conv_s2_all(w_hwio, out=None):
    """the four stride-2 data-gradient operands (dafk_pack_conv mode 2) back to back in class order 2*pa + pb"""
    KH, KW, Cin, Cout = w_hwio.shape
    Cip, Cop = (Cin + 63) // 64 * 64, (Cout + 63) // 64 * 64
    n = (KH // 2) * (KW // 2) * Cip * Cop
    if out is None:
        out = torch.empty(4 * n, dtype=torch.bfloat16, device=w_hwio.device)
    for pa in (0, 1):
        for pb in (0, 1):
            pack_conv(w_hwio, 2, pa, pb, out=out[(2 * pa + pb) * n:(2 * pa + pb + 1) * n])
    return out


def conv_tc_dgrad_s2(dy, wp4, x_shape, Cin, KH, KW, out_dtype=torch.float32, row_off=0, rows_per_tap=None):
    """dx of a valid stride-2 convolution with an even kernel, all four parity classes in one launch"""
    _chk(dy, wp4)
    N, H, W, _ = x_shape
    _, Ho, Wo, Cout = dy.shape
    rpt = rows_per_tap if rows_per_tap is not None else (Cin + 63) // 64 * 64
    dx = torch.empty((N, H, W, Cin), dtype=out_dtype, device=dy.device)
    flops = 2.0 * N * Ho * Wo * KH * KW * Cin * Cout
    nb = dy.numel() * 2.0 + dx.numel() * dx.element_size()
    instrument.timed("conv_tc_fwd+dgrad (tcgen05)", flops, nb,
                     lambda: call("conv_tc_dgrad_s2", dy, Cout, wp4, rpt, row_off, dx, _dt(dx), N, Ho, Wo, Cin, KH, KW, H, W, _S()),
                     tag=(N, Ho, Wo, Cout, Cin, KH, "s2-dgrad"))
    return dx


def _wgrad_halo(Cin, Cout, KH, KW, stride, pad):
    if not (KH == 3 and KW == 3 and stride == 1 and pad == 1 and Cin % 64 == 0 and Cout % 64 == 0):
        return False
    if WGRAD_HALO in ("0", "1"):
        return WGRAD_HALO == "1"
    # measured on B200 (profiles/r1_bench_tc.txt): wins while a pixel carries few channels (the per-tap kernel is
    # L2-bound there); the 128x128-tile per-tap kernel is faster from 256 channels on
    return Cin <= 128 and Cout <= 128


def conv3x3_tc_wgrad_halo(x, dy, dw, cin_off=0):
    _chk(x, dy, dw)
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    fl = 2.0 * N * H * W * Cout * 9 * Cin
    instrument.timed("conv_tc_wgrad (tcgen05)", fl, 2.0 * (x.numel() + dy.numel()),
                     lambda: call("conv3x3_tc_wgrad_halo", x, Cin, cin_off, dw.shape[2], dy, Cout, dw, N, H, W, _S()),
                     tag=(N, H, W, Cin, Cout, 3, 1, "halo"))


def conv_tc_wgrad(x, dy, dw, cin_off, KH, KW, stride, pad):
    _chk(x, dy, dw)
    N, H, W, Cin = x.shape
    _, Ho, Wo, Cout = dy.shape
    if _wgrad_halo(Cin, Cout, KH, KW, stride, pad):
        return conv3x3_tc_wgrad_halo(x, dy, dw, cin_off)
    fl = 2.0 * N * Ho * Wo * Cout * KH * KW * Cin
    instrument.timed("conv_tc_wgrad (tcgen05)", fl, 2.0 * (x.numel() + dy.numel()),
                     lambda: call("conv_tc_wgrad", x, Cin, cin_off, dw.shape[2], dy, Cout, dw, N, H, W, KH, KW, stride,
                                  pad, Ho, Wo, _S()), tag=(N, H, W, Cin, Cout, KH, stride))


def conv3x3_tc_wgrad(x, dy, dw, cin_off=0):
    """dw[3,3,cin_total,Cout] (f32 HWIO) += x (*) dy for the channel block starting at cin_off."""
    _chk(x, dy, dw)
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    fl = 2.0 * N * H * W * Cout * 9 * Cin
    instrument.timed("conv_tc_wgrad (tcgen05)", fl, 2.0 * (x.numel() + dy.numel()),
                     lambda: call("conv3x3_tc_wgrad", x, Cin, cin_off, dw.shape[2], dy, Cout, dw, N, H, W, _S()),
                     tag=(N, H, W, Cin, Cout, 3, 1))


# ---------------------------------------------------------------------------- narrow-channel tcgen05 convolution
USE_NC = True        # tests switch it off to exercise the CUDA-core kernels


def nc_supported(Cin, Cout, KH, KW, W, pad, kind):
    """kind 0 forward, 1 stride-1 data gradient (pass the LAYER's Cin/Cout/pad), 2 weight gradient"""
    if not USE_NC:
        return False
    f = _lib.lib().fn["dafk_conv_nc_supported"]
    if kind == 1:
        return bool(f(Cout, Cin, KH, KW, W + 2 * pad - KW + 1, KH - 1 - pad, 0))
    return bool(f(Cin, Cout, KH, KW, W, pad, kind))


def nc_wgrad_stages_raw(x_shape, x_dtype, Cout, k, pad, dy_dtype=torch.bfloat16):
    """would conv_nc_wgrad bring rows of these dtypes in with bulk copies (fast on a bf16 output gradient)?"""
    if not USE_NC:
        return False
    N, H, W, Cin = x_shape
    code = {torch.float32: DAFK_F32, torch.bfloat16: DAFK_BF16}
    return bool(_lib.lib().fn["dafk_conv_nc_wgrad_stages_raw"](N, H, W, Cin, Cout, k, k, pad, code[x_dtype], code[dy_dtype]))


def pack_conv_nc(w_hwio, mode, out=None):
    """mode 0: forward operand; mode 1: stride-1 data-gradient operand (mirrored taps, transposed channels)"""
    _chk(w_hwio)
    KH, KW, Cin, Cout = w_hwio.shape
    ck, nk = (Cin, Cout) if mode == 0 else (Cout, Cin)
    n = _lib.lib().fn["dafk_conv_nc_packed_elems"](ck, nk, KH, KW)
    wp = torch.empty(n, dtype=torch.bfloat16, device=w_hwio.device) if out is None else out
    call("pack_conv_nc", w_hwio, wp, KH, KW, Cin, Cout, mode, _S())
    return wp


def pack_conv_nc_scaled(w_hwio, scale, out=None):
    _chk(w_hwio, scale)
    KH, KW, Cin, Cout = w_hwio.shape
    n = _lib.lib().fn["dafk_conv_nc_packed_elems"](Cin, Cout, KH, KW)
    wp = torch.empty(n, dtype=torch.bfloat16, device=w_hwio.device) if out is None else out
    call("pack_conv_nc_scaled", w_hwio, scale, wp, KH, KW, Cin, Cout, _S())
    return wp


def conv_nc_fwd(x, wp, bias, Cout, KH, KW, pad, act=ACT_NONE, alpha=0.0, out_dtype=torch.float32):
    """y = act(conv(x, w) + bias), stride 1; x f32/bf16 NHWC; wp from pack_conv_nc"""
    _chk(x, wp, bias)
    N, H, W, Cin = x.shape
    Ho, Wo = H + 2 * pad - KH + 1, W + 2 * pad - KW + 1
    y = torch.empty((N, Ho, Wo, Cout), dtype=out_dtype, device=x.device)
    fl = 2.0 * N * Ho * Wo * Cout * KH * KW * Cin
    nb = x.numel() * x.element_size() + y.numel() * y.element_size()
    instrument.timed("conv_nc_fwd+dgrad (tcgen05)", fl, nb,
                     lambda: call("conv_nc_fwd", x, _dt(x), wp, bias, y, _dt(y), N, H, W, Cin, Cout, KH, KW, pad, act,
                                  float(alpha), _S()), tag=(N, H, W, Cin, Cout, KH, pad, str(x.dtype)[6:], str(out_dtype)[6:]))
    return y


def conv_nc_fwd_cat(xa, xb, wp, bias, Cout, KH, KW, pad, act=ACT_NONE, alpha=0.0, out_dtype=torch.float32):
    """conv_nc_fwd on concat([xa, xb], -1) read where the sources lie (same dtype, xa a multiple of 8 channels)"""
    _chk(xa, xb, wp, bias)
    N, H, W, Ca = xa.shape
    Cb = xb.shape[-1]
    assert xa.dtype == xb.dtype and tuple(xb.shape[:3]) == (N, H, W) and Ca % 8 == 0
    Ho, Wo = H + 2 * pad - KH + 1, W + 2 * pad - KW + 1
    y = torch.empty((N, Ho, Wo, Cout), dtype=out_dtype, device=xa.device)
    fl = 2.0 * N * Ho * Wo * Cout * KH * KW * (Ca + Cb)
    nb = (xa.numel() + xb.numel()) * xa.element_size() + y.numel() * y.element_size()
    instrument.timed("conv_nc_fwd+dgrad (tcgen05)", fl, nb,
                     lambda: call("conv_nc_fwd_cat", xa, Ca, xb, Cb, _dt(xa), wp, bias, y, _dt(y), N, H, W, Cout, KH, KW, pad, act,
                                  float(alpha), _S()),
                     tag=(N, H, W, Ca + Cb, Cout, KH, pad, str(xa.dtype)[6:], str(out_dtype)[6:], "cat"))
    return y


def conv_nc_wgrad_cat(xa, xb, dy, dw, db, pad):
    """conv_nc_wgrad with x = concat([xa, xb], -1) read where the sources lie"""
    _chk(xa, xb, dy, dw, db)
    N, H, W, Ca = xa.shape
    Cb = xb.shape[-1]
    KH, KW, Cin, Cout = dw.shape
    assert Cin == Ca + Cb and xa.dtype == xb.dtype
    fl = 2.0 * dy.numel() * KH * KW * Cin
    nb = (xa.numel() + xb.numel()) * xa.element_size() + dy.numel() * dy.element_size()
    instrument.timed("conv_nc_wgrad (tcgen05)", fl, nb,
                     lambda: call("conv_nc_wgrad_cat", xa, Ca, xb, Cb, _dt(xa), dy, _dt(dy), dw, db, N, H, W, Cout, KH, KW, pad,
                                  _S()),
                     tag=(N, H, W, Cin, Cout, KH, pad, str(xa.dtype)[6:], str(dy.dtype)[6:], "cat"))


def conv_nc_wgrad(x, dy, dw, db, pad):
    """dw[KH,KW,Cin,Cout] += x (*) dy; db += sum dy (db may be None); stride 1"""
    _chk(x, dy, dw, db)
    N, H, W, Cin = x.shape
    KH, KW, _, Cout = dw.shape
    fl = 2.0 * dy.numel() * KH * KW * Cin
    nb = x.numel() * x.element_size() + dy.numel() * dy.element_size()
    instrument.timed("conv_nc_wgrad (tcgen05)", fl, nb,
                     lambda: call("conv_nc_wgrad", x, _dt(x), dy, _dt(dy), dw, db, N, H, W, Cin, Cout, KH, KW, pad, _S()),
                     tag=(N, H, W, Cin, Cout, KH, pad, str(x.dtype)[6:], str(dy.dtype)[6:]))


# ---------------------------------------------------------------------------- pointwise 64 -> <=8 heads
def conv1x1_supported(Cin, Cout):
    return USE_NC and bool(_lib.lib().fn["dafk_conv1x1_supported"](Cin, Cout))


def conv1x1_fwd(x, w, bias, round_bf16=True):
    """x bf16 [..,64], w f32 [1,1,64,Cout] -> y f32 [..,Cout]"""
    _chk(x, w, bias)
    assert x.dtype == torch.bfloat16
    Cin, Cout = w.shape[-2], w.shape[-1]
    M = x.numel() // Cin
    y = f32(*x.shape[:-1], Cout)
    instrument.timed("conv1x1 (64 -> <=8 heads)", 2.0 * M * Cin * Cout, 2.0 * x.numel() + 4.0 * y.numel(),
                     lambda: call("conv1x1_fwd", x, w, bias, y, M, Cin, Cout, int(round_bf16), _S()), tag=("fwd", M, Cout))
    return y


def conv1x1_dgrad(dy, w, round_bf16=True):
    """dy f32 [..,Cout] -> dx bf16 [..,64]"""
    _chk(dy, w)
    assert dy.dtype == torch.float32
    Cin, Cout = w.shape[-2], w.shape[-1]
    M = dy.numel() // Cout
    dx = torch.empty(tuple(dy.shape[:-1]) + (Cin,), dtype=torch.bfloat16, device=dy.device)
    instrument.timed("conv1x1 (64 -> <=8 heads)", 2.0 * M * Cin * Cout, 4.0 * dy.numel() + 2.0 * dx.numel(),
                     lambda: call("conv1x1_dgrad", dy, w, dx, M, Cin, Cout, int(round_bf16), _S()), tag=("dgrad", M, Cout))
    return dx


def conv1x1_wgrad(x, dy, dw, db, round_bf16=True):
    """dw[1,1,64,Cout] += x^T dy; db += sum dy (db may be None)"""
    _chk(x, dy, dw, db)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.float32
    Cin, Cout = dw.shape[-2], dw.shape[-1]
    M = x.numel() // Cin
    instrument.timed("conv1x1 (64 -> <=8 heads)", 2.0 * M * Cin * Cout, 2.0 * x.numel() + 4.0 * dy.numel(),
                     lambda: call("conv1x1_wgrad", x, dy, dw, db, M, Cin, Cout, int(round_bf16), _S()), tag=("wgrad", M, Cout))


def space_to_depth2(x):
    """[N,H,W,C] -> bf16 [N,ceil(H/2),ceil(W/2),4C] (zero beyond odd sizes)"""
    _chk(x)
    N, H, W, C = x.shape
    y = torch.empty((N, (H + 1) // 2, (W + 1) // 2, 4 * C), dtype=torch.bfloat16, device=x.device)
    call("space_to_depth2", x, _dt(x), y, N, H, W, C, _S())
    return y


def space_to_depth2_cat(xa, xb):
    """space_to_depth2(concat([xa, xb], -1)) without the concatenated copy"""
    _chk(xa, xb)
    N, H, W, Ca = xa.shape
    Cb = xb.shape[-1]
    assert tuple(xb.shape[:3]) == (N, H, W)
    y = torch.empty((N, (H + 1) // 2, (W + 1) // 2, 4 * (Ca + Cb)), dtype=torch.bfloat16, device=xa.device)
    call("space_to_depth2_cat", xa, _dt(xa), Ca, xb, _dt(xb), Cb, y, N, H, W, _S())
    return y


def depth_to_space2_split(y, H, W, Ca, Cb, want_a=True, want_b=True):
    """inverse of space_to_depth2_cat for gradients: fp32 [N,H,W,Ca] and [N,H,W,Cb] (None where not wanted)"""
    _chk(y)
    N = y.shape[0]
    assert y.shape[-1] == 4 * (Ca + Cb)
    ga = torch.empty((N, H, W, Ca), dtype=torch.float32, device=y.device) if want_a else None
    gb = torch.empty((N, H, W, Cb), dtype=torch.float32, device=y.device) if want_b else None
    call("depth_to_space2_split", y, _dt(y), ga, Ca, gb, Cb, N, H, W, _S())
    return ga, gb


def depth_to_space2(y, H, W, out_dtype=torch.float32):
    _chk(y)
    N, _, _, C4 = y.shape
    x = torch.empty((N, H, W, C4 // 4), dtype=out_dtype, device=y.device)
    call("depth_to_space2", y, _dt(y), x, _dt(x), N, H, W, C4 // 4, _S())
    return x


def conv_s2d_weights(w, out=None):
    """HWIO [KH,KW,C,Co] -> [ceil(KH/2),ceil(KW/2),4C,Co]"""
    _chk(w)
    KH, KW, C, Co = w.shape
    w2 = torch.empty(((KH + 1) // 2, (KW + 1) // 2, 4 * C, Co), dtype=torch.float32, device=w.device) if out is None else out
    call("conv_s2d_weights", w, w2, KH, KW, C, Co, 0, _S())
    return w2


def conv_s2d_weights_bwd_(dw, dw2):
    """dw += rearranged^T(dw2)"""
    _chk(dw, dw2)
    KH, KW, C, Co = dw.shape
    call("conv_s2d_weights", dw, dw2, KH, KW, C, Co, 1, _S())


# ---------------------------------------------------------------------------- dense
def dense_fwd(x, w, bias):
    _chk(x, w, bias)
    B, K = x.shape
    N = w.shape[1]
    y = f32(B, N)
    call("dense_fwd", x, w, bias, y, B, K, N, _S())
    return y


def dense_bwd_data(dy, w):
    _chk(dy, w)
    B, N = dy.shape
    K = w.shape[0]
    dx = f32(B, K)
    call("dense_bwd_data", dy, w, dx, B, K, N, _S())
    return dx


def dense_bwd_weight(x, dy, dw, db):
    _chk(x, dy, dw, db)
    B, K = x.shape
    call("dense_bwd_weight", x, dy, dw, db, B, K, dy.shape[1], _S())


# ---------------------------------------------------------------------------- TPS
_tps_consts_cache = {}


def tps_consts(cp_h, cp_w, device):
    key = (cp_h, cp_w, str(device))
    if key not in _tps_consts_cache:
        n = cp_h * cp_w
        nfl = _lib.lib().fn["dafk_tps_consts_floats"](n)
        host = torch.empty(nfl, dtype=torch.float32)
        call("tps_build_constants", cp_h, cp_w, host)
        _tps_consts_cache[key] = host.to(device)
    return _tps_consts_cache[key]


_tps_phi_cache = {}
TPS_PHI_TABLE = os.environ.get("DAFK_TPS_PHI_TABLE", "1") != "0"


def tps_phi_table(H, W, cp, device):
    """phi(|q - c|^2) for every pixel and control point of one geometry: built once, then read by every warp"""
    key = (H, W, cp[0], cp[1], str(device))
    if key not in _tps_phi_cache:
        n = cp[0] * cp[1]
        tab = f32(int(_lib.lib().fn["dafk_tps_phi_table_floats"](H, W, n)), device=device)
        call("tps_phi_table", tps_consts(cp[0], cp[1], device), tab, H, W, n, _S())
        _tps_phi_cache[key] = tab
    return _tps_phi_cache[key]


def tps_warp_fwd(vol, theta, cp=(5, 5), want_locs=False):
    _chk(vol, theta)
    B, H, W, C = vol.shape
    consts = tps_consts(cp[0], cp[1], vol.device)
    out = torch.empty_like(vol)
    locs = f32(B, H * W, 2) if want_locs else None
    if TPS_PHI_TABLE:
        tab = tps_phi_table(H, W, cp, vol.device)
        coef_ws = f32(B * (cp[0] * cp[1] + 3) * 2, device=vol.device)
        instrument.timed("tps_warp_fwd", 0, 8.0 * vol.numel(),
                         lambda: call("tps_warp_fwd_tab", vol, theta, consts, tab, coef_ws, out, locs, B, H, W, C,
                                      cp[0] * cp[1], _S()))
        return out, locs
    instrument.timed("tps_warp_fwd", 0, 8.0 * vol.numel(),
                     lambda: call("tps_warp_fwd", vol, theta, consts, out, locs, B, H, W, C, cp[0] * cp[1], _S()))
    return out, locs


def tps_warp_bwd(vol, theta, dout, cp=(5, 5), need_dvol=True):
    _chk(vol, theta, dout)
    B, H, W, C = vol.shape
    n = cp[0] * cp[1]
    consts = tps_consts(cp[0], cp[1], vol.device)
    dvol = None
    if need_dvol:
        dvol = zero_(torch.empty_like(vol))
    dtheta = f32(B, n, 2)
    ws = torch.empty(B * (n + 3) * 2, dtype=torch.float64, device=vol.device)
    if TPS_PHI_TABLE:
        tab = tps_phi_table(H, W, cp, vol.device)
        instrument.timed("tps_warp_bwd", 0, 12.0 * vol.numel(),
                         lambda: call("tps_warp_bwd_tab", vol, theta, consts, tab, dout, dvol, dtheta, ws, B, H, W, C, n, _S()))
        return dvol, dtheta
    instrument.timed("tps_warp_bwd", 0, 12.0 * vol.numel(),
                     lambda: call("tps_warp_bwd", vol, theta, consts, dout, dvol, dtheta, ws, B, H, W, C, n, _S()))
    return dvol, dtheta


def tps_solve(train_points, train_values, order=2, reg=0.0):
    _chk(train_points, train_values)
    B, n, _ = train_points.shape
    k = train_values.shape[-1]
    w, v = f32(B, n, k), f32(B, 3, k)
    call("tps_solve_batched", train_points, train_values, w, v, B, n, k, order, float(reg), _S())
    return w, v


def tps_apply(query, train_points, w, v, order=2):
    _chk(query, train_points, w, v)
    B, n, k = w.shape
    qb = 1 if (query.shape[0] == B and B > 1) else 0
    m = query.shape[1]
    out = f32(B, m, k)
    call("tps_apply", query, train_points, w, v, out, B, m, n, k, order, qb, _S())
    return out


def resampler_fwd(vol, warp):
    _chk(vol, warp)
    B, H, W, C = vol.shape
    m = warp.shape[1]
    out = f32(B, m, C)
    call("resampler_fwd", vol, warp, out, B, H, W, C, m, _S())
    return out


# ---------------------------------------------------------------------------- losses
def segloss(pred, target, nch, use_bce, weight, loss, lambda_bce=0.01, want_grad=True):
    """loss[0] += weight*(dice + lambda*wbce); returns d(weight*loss)/dpred (or None)."""
    _chk(pred, target, loss)
    B, Cp, Ct = pred.shape[0], pred.shape[-1], target.shape[-1]
    HW = pred.numel() // (B * Cp)
    nws = _lib.lib().fn["dafk_segloss_ws_doubles"](B, Cp)
    ws = torch.empty(nws, dtype=torch.float64, device=pred.device)
    call("segloss_fwd", pred, Cp, target, Ct, nch, int(use_bce), float(lambda_bce), ws, B, HW, _S())
    call("segloss_finish", ws, float(weight), loss, B, Cp, nch, int(use_bce), float(lambda_bce), HW, _S())
    if not want_grad:
        return None
    dpred = torch.empty_like(pred)
    call("segloss_bwd", pred, Cp, target, Ct, nch, int(use_bce), float(lambda_bce), ws, float(weight), dpred, B, HW, _S())
    return dpred


def l1l2_loss(pred, target, kind, weight, loss, cval=0.0, want_grad=True):
    _chk(pred, target, loss)
    dpred = torch.empty_like(pred) if want_grad else None
    call("l1l2_loss", pred, target, float(cval), kind, float(weight), loss, dpred, pred.numel(), _S())
    return dpred


def vae_fwd(mu, logvar, eps, weight, loss):
    _chk(mu, logvar, eps)
    B, Z = mu.shape
    z, klv = torch.empty_like(mu), f32(B, 1)
    call("vae_fwd", mu, logvar, eps, z, klv, float(weight), loss, B, Z, _S())
    return z, klv


def vae_bwd(mu, logvar, eps, dz, weight):
    B, Z = mu.shape
    dmu, dlv = torch.empty_like(mu), torch.empty_like(mu)
    call("vae_bwd", mu, logvar, eps, dz, float(weight), dmu, dlv, B, Z, _S())
    return dmu, dlv


def spectral_reg(W2d, u0, alpha, loss, dW):
    _chk(W2d, u0)
    dim, cout = W2d.shape
    ws = f32(2 * dim + cout + 4)
    call("spectral_reg", W2d, u0, float(alpha), loss, dW, ws, dim, cout, _S())


def adam_step(p, g, m, v, shadow, lr_t, b1=0.9, b2=0.999, eps=1e-7, grad_scale=1.0):
    _chk(p, g, m, v)
    call("adam_step", p, g, m, v, shadow, p.numel(), float(lr_t), float(b1), float(b2), float(eps), float(grad_scale), _S())


def adam_tick(state, lr, b1=0.9, b2=0.999):
    _chk(state)
    call("adam_tick", state, float(lr), float(b1), float(b2), _S())


def adam_step_dev(p, g, m, v, shadow, state, b1=0.9, b2=0.999, eps=1e-7, grad_scale=1.0):
    _chk(p, g, m, v, state)
    call("adam_step_dev", p, g, m, v, shadow, p.numel(), state, float(b1), float(b2), float(eps), float(grad_scale), _S())


# ---------------------------------------------------------------------------- SPADE / balancer
def in_stats(x):
    _chk(x)
    B = x.shape[0]
    acc = torch.empty(2 * B, dtype=torch.float64, device=x.device)
    zero_(acc)
    call("in_stats", x, acc, B, x.numel() // B, _S())
    return acc


def spade_fwd(x, acc, gamma, beta, act=ACT_LRELU, alpha=0.2, eps=1e-3):
    _chk(x, gamma, beta)
    B = x.shape[0]
    y = torch.empty_like(x)
    call("spade_fwd", x, acc, gamma, beta, y, B, x.numel() // B, float(eps), act, float(alpha), _S())
    return y


def spade_bwd(dy, x, acc, gamma, beta, act=ACT_LRELU, alpha=0.2, eps=1e-3):
    _chk(dy, x, gamma, beta)
    B = x.shape[0]
    dx, dg, db = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    ws = torch.empty(2 * B, dtype=torch.float64, device=x.device)
    call("spade_bwd", dy, x, acc, gamma, beta, dx, dg, db, ws, B, x.numel() // B, float(eps), act, float(alpha), _S())
    return dx, dg, db


def spade_cond_fwd(x, gamma, beta):
    """layers/spade.py:41-58: x*(1+gamma)+beta"""
    _chk(x, gamma, beta)
    assert x.shape == gamma.shape == beta.shape and x.dtype == torch.float32
    y = torch.empty_like(x)
    call("spade_cond_fwd", x, gamma, beta, y, x.numel(), _S())
    return y


def spade_cond_bwd(dy, x, gamma):
    _chk(dy, x, gamma)
    dx, dg = torch.empty_like(x), torch.empty_like(x)
    call("spade_cond_bwd", dy, x, gamma, dx, dg, x.numel(), _S())
    return dx, dg


def in_affine_fwd(x, acc, gamma, beta, act=ACT_NONE, alpha=0.0, eps=1e-3):
    """keras_contrib InstanceNormalization(axis=None) with scalar gamma/beta + activation"""
    _chk(x, gamma, beta)
    B = x.shape[0]
    y = torch.empty_like(x)
    call("in_affine_fwd", x, acc, gamma, beta, y, B, x.numel() // B, float(eps), act, float(alpha), _S())
    return y


def in_affine_bwd(dy, x, acc, gamma, beta, dgamma, dbeta, act=ACT_NONE, alpha=0.0, eps=1e-3):
    """returns dx; accumulates the scalar parameter gradients into dgamma / dbeta (None = frozen)"""
    _chk(dy, x, gamma, beta)
    B = x.shape[0]
    dx = torch.empty_like(x)
    ws = torch.empty(2 * B + 2, dtype=torch.float64, device=x.device)
    call("in_affine_bwd", dy, x, acc, gamma, beta, dx, dgamma, dbeta, ws, B, x.numel() // B, float(eps), act, float(alpha), _S())
    return dx


def pair_dice(a, b, want_ws=False):
    _chk(a, b)
    B = a.shape[0]
    out = f32(B, 1)
    ws = torch.empty(3 * B, dtype=torch.float64, device=a.device)
    call("pair_dice", a, b, out, ws, B, a.numel() // B, _S())
    return (out, ws) if want_ws else out


def pair_dice_bwd(a, b, ws, g, need_a=True, need_b=True):
    """g [B,1] = d loss / d dice -> (da, db)"""
    _chk(a, b, ws, g)
    B = a.shape[0]
    da = torch.empty_like(a) if need_a else None
    db = torch.empty_like(b) if need_b else None
    call("pair_dice_bwd", a, b, ws, g, da, db, B, a.numel() // B, _S())
    return da, db


# ---------------------------------------------------------------------------- augmentation
def rotate_bilinear(x, theta):
    """keras ImageDataGenerator(rotation_range) on a staged batch: x f32 [B,H,W,C], theta [B] radians"""
    _chk(x, theta)
    B, H, W, C = x.shape
    y = torch.empty_like(x)
    call("rotate_bilinear", x, theta, y, B, H, W, C, _S())
    return y


def mask_residual_(m):
    """rebuild the background channel (last) of a staged mask batch from its other channels, in place"""
    _chk(m)
    C = m.shape[-1]
    call("mask_residual", m, m.numel() // C, C, _S())
    return m


# ---------------------------------------------------------------------------- automated-pairing losses
def segloss_pb_fwd(pred, target, nch, L_row, lambda_bce=0.01):
    """per-sample dice + lambda * swapped per-batch wBCE (costs.py:88-108,138-143) -> L_row[B]; returns the workspace"""
    _chk(pred, target, L_row)
    B, C = pred.shape[0], pred.shape[-1]
    assert target.shape[-1] == C
    HW = pred.numel() // (B * C)
    n = int(_lib.lib().fn["dafk_segloss_pb_ws_doubles"](B, C))
    ws = torch.empty(n, dtype=torch.float64, device=pred.device)
    call("segloss_pb_fwd", pred, target, C, nch, float(lambda_bce), ws, L_row, B, HW, _S())
    return ws


def segloss_pb_bwd(pred, target, nch, ws, coef_row, lambda_bce=0.01):
    _chk(pred, target, ws, coef_row)
    B, C = pred.shape[0], pred.shape[-1]
    HW = pred.numel() // (B * C)
    g = torch.empty_like(pred)
    call("segloss_pb_bwd", target, C, nch, float(lambda_bce), ws, coef_row, g, B, HW, _S())
    return g


def mae_pb_fwd(pred, target, L_row):
    _chk(pred, target, L_row)
    B = pred.shape[0]
    ws = torch.empty(B, dtype=torch.float64, device=pred.device)
    call("mae_pb_fwd", pred, target, ws, L_row, B, pred.numel() // B, _S())
    return ws


def mae_pb_bwd(pred, target, coef_row):
    _chk(pred, target, coef_row)
    B = pred.shape[0]
    g = torch.empty_like(pred)
    call("mae_pb_bwd", pred, target, coef_row, g, B, pred.numel() // B, _S())
    return g


def pair_combine(w, L, weight, loss, want_dw=True):
    """w [B,P] (or None = ones), L [P,B] -> (dw [B,P] or None, coef [P,B]); loss[0] += weight/B * sum w*L"""
    _chk(w, L, loss)
    P, B = L.shape
    dw = torch.empty((B, P), dtype=torch.float32, device=L.device) if (want_dw and w is not None) else None
    coef = torch.empty_like(L)
    call("pair_combine", w, L, float(weight), loss, dw, coef, B, P, _S())
    return dw, coef

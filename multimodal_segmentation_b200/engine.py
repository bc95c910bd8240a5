"""Host-side execution engine: parameter arenas, a minimal reverse-mode tape and the layer
primitives (Conv2D, BatchNormalization, Dense, ...) that the model components are built from.

Design (B200-first, not a Keras port):
  * parameters of all components of a model live in ONE flat fp32 arena with a same-shaped flat
    gradient arena -> one fused Adam launch per optimizer step and one NCCL all-reduce bucket;
  * every arithmetic op is a hand-written kernel reached through ``ops`` (ctypes -> libdafk.so);
    torch only owns the device memory;
  * the backward pass is a list of closures recorded while the forward runs (``Tape``), so a
    component that is called several times in one graph (Segmentor x4, Decoder x6 in
    models/dafnet.py:163-222) simply records several nodes and accumulates into the same
    parameter gradients;
  * wide 3x3 convolutions run on tcgen05 with bf16 operands: the producing kernel (BN-apply,
    pooling, upsampling) writes bf16 directly, gradients of bf16 maps are bf16.
"""
import math
import os

import numpy as np
import torch

from . import ops
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH

# Use the tcgen05 tensor-core path for eligible 3x3 convolutions (bf16 operands, fp32 accumulate).
# Parity tests against the fp32 oracle at 1e-4 switch it off.
USE_TC = True

# Store the outputs of convolutions that feed a BatchNorm in the feature dtype (bf16 on the tensor-core path) instead of
# fp32: halves the bytes of the largest tensors of the step (-3 % step time).  Costs one more bf16 rounding per layer
# (what mixed-precision training does everywhere); tests/test_models_gpu.py bounds the effect with both settings.
RAW_BF16 = True

# Keep the FiLM decoder's 8-channel activations (and their gradients) in bf16: halves the HBM bytes of its convolutions and
# element-wise passes and lets the 8 -> 8 convolutions stage rows with cp.async.bulk (csrc/conv_nc.cu, DAFK_NC_BULK=1).
# Opt-in (DAFK_DEC_BF16=1) until its effect on the reconstruction losses has been measured on the full step.
DEC_BF16 = os.environ.get("DAFK_DEC_BF16", "0") == "1"

ACT = {None: ACT_NONE, "linear": ACT_NONE, "relu": ACT_RELU, "lrelu": ACT_LRELU, "tanh": ACT_TANH}


def dec_dtype():
    """storage dtype of the FiLM decoder's feature maps"""
    return torch.bfloat16 if (USE_TC and RAW_BF16 and DEC_BF16) else torch.float32


def feat_dtype():
    return torch.bfloat16 if USE_TC else torch.float32


# --------------------------------------------------------------------------------------------
# tape
# --------------------------------------------------------------------------------------------
class Var:
    """A value in the forward graph.  ``data`` is a CUDA tensor (f32 or bf16).

    Gradients: the first contribution is kept as it arrives; a second one is held back in ``_grad2`` instead of being
    added at once, so that a consumer that applies an activation backward next can sum the two inside that pass
    (``take_grad_pair``); reading ``grad`` materialises the sum."""
    __slots__ = ("data", "_grad", "_grad2", "requires_grad", "grad_dtype", "grad_owned", "bias_sink", "bn_acc")

    def __init__(self, data, requires_grad=False, grad_dtype=None):
        self.data = data
        self._grad = None
        self._grad2 = None
        self.requires_grad = requires_grad
        self.grad_dtype = grad_dtype if grad_dtype is not None else data.dtype
        self.grad_owned = False
        self.bias_sink = None
        self.bn_acc = None       # fp64 [2C] sum / sum of squares of `data`, when the producing convolution took them

    @property
    def grad(self):
        if self._grad2 is not None:
            g2, self._grad2 = self._grad2, None
            if self.grad_owned:
                ops.add_(self._grad, g2)
            else:
                self._grad = ops.add(self._grad, g2)
                self.grad_owned = True
        return self._grad

    @grad.setter
    def grad(self, g):
        self._grad = g
        self._grad2 = None

    def take_grad_pair(self):
        """(g1, g2 or None, g1 is ours to overwrite) and clear: for consumers that fuse the sum into their own first pass"""
        g1, g2, owned = self._grad, self._grad2, self.grad_owned
        self._grad = self._grad2 = None
        return g1, g2, owned

    @property
    def shape(self):
        return self.data.shape


class Ctx:
    """Execution context of one forward pass: ``tape`` (None = no gradient) and the Keras
    learning phase (``training`` selects batch statistics in BatchNormalization)."""

    def __init__(self, tape=None, training=False):
        self.tape = tape
        self.training = training

    def rec(self, *inputs):
        if self.tape is None:
            return False
        # the parameters a node is about to be recorded for: Tape.record() attaches them to the node, so that the data-parallel
        # trainer knows after which node of the backward pass a range of the gradient arena is final (Tape.last_writers)
        self.tape._pending = [v for v in inputs if isinstance(v, Param) and v.requires_grad]
        return any(v.requires_grad for v in inputs if v is not None)


class Tape:
    def __init__(self):
        self.nodes = []
        self.node_params = []
        self._pending = []

    def record(self, fn):
        self.nodes.append(fn)
        self.node_params.append(self._pending)
        self._pending = []

    def backward(self, after_node=None):
        """run the recorded nodes in reverse; ``after_node(i)`` is called when node i (forward order) has run"""
        for i in range(len(self.nodes) - 1, -1, -1):
            self.nodes[i]()
            if after_node is not None:
                after_node(i)
        self.nodes = []
        self.node_params = []

    def last_writers(self, ranges):
        """for every (arena, a, b) range of a gradient arena: the forward index of the FIRST node that touches a parameter
        inside it -- the backward pass runs the nodes in reverse, so once that node has run the range is final.
        None: no recorded node writes there (frozen or unused parameters)."""
        first = [None] * len(ranges)
        for i, ps in enumerate(self.node_params):
            for p in ps:
                lo, hi = p.offset, p.offset + p.size
                for k, (arena, a, b) in enumerate(ranges):
                    if first[k] is None and p.arena is arena and lo < b and hi > a:
                        first[k] = i
        return first


def accumulate(var, g, owned=True):
    """var.grad += g.  ``owned`` = the caller hands over the buffer (it may be updated in place later)."""
    if var is None or not var.requires_grad:
        return
    if g.dtype != var.grad_dtype:
        g = ops.cast(g, var.grad_dtype)
        owned = True
    if var._grad is None:
        var.grad = g
        var.grad_owned = owned
    elif var._grad2 is None:
        var._grad2 = g               # summed by whoever reads the gradient (possibly inside its activation backward)
    elif var.grad_owned:
        ops.add_(var._grad, g)
    else:
        var._grad = ops.add(var._grad, g)
        var.grad_owned = True


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
class Param:
    """A trainable tensor: views into the flat parameter / gradient arenas."""

    def __init__(self, name, shape, init):
        self.name = name
        self.shape = tuple(int(s) for s in shape)
        self.size = int(np.prod(self.shape))
        self.init = np.ascontiguousarray(init, dtype=np.float32).reshape(self.shape)
        self.arena = None
        self.offset = None
        self.data = None
        self.grad = None
        self.requires_grad = True      # toggled by make_trainable (utils/sdnet_utils.py:40-53)

    def numpy(self):
        return self.data.detach().cpu().numpy().copy()


class Arena:
    """Flat fp32 storage for a group of parameters (+ a gradient arena of the same shape) or for
    non-trainable state such as BatchNorm moving statistics (``with_grad=False``)."""

    def __init__(self, with_grad=True):
        self.params = []
        self.with_grad = with_grad
        self.flat = None
        self.gflat = None
        self.version = 0           # bumped by every optimizer step that touches the arena
        self._cursor = 0

    def add(self, name, shape, init):
        p = Param(name, shape, init)
        p.arena = self
        p.offset = self._cursor
        self._cursor += (p.size + 3) // 4 * 4     # keep every tensor 16-byte aligned
        self.params.append(p)
        return p

    def to_device(self, device="cuda"):
        host = np.zeros(max(self._cursor, 4), np.float32)
        for p in self.params:
            host[p.offset:p.offset + p.size] = p.init.ravel()
        if not torch.cuda.is_available():
            # host-only mode (CPU test box): weights can be built, inspected and exported, but nothing can
            # be computed -- every kernel wrapper raises on a non-CUDA tensor.
            self.flat = torch.from_numpy(host)
            self.gflat = None
            for p in self.params:
                p.data = self.flat[p.offset:p.offset + p.size].view(p.shape)
                p.grad = None
            return self
        self.flat = torch.from_numpy(host).to(device)
        self.gflat = ops.zeros(self.flat.numel()) if self.with_grad else None
        for p in self.params:
            p.data = self.flat[p.offset:p.offset + p.size].view(p.shape)
            p.grad = self.gflat[p.offset:p.offset + p.size].view(p.shape) if self.with_grad else None
        return self

    def numel(self):
        return self._cursor


class Adam:
    """Keras 2.1.6 Adam (models/dafnet.py:155 etc.): one independent state per compiled trainer.
    Works on merged contiguous ranges of the arenas, one fused launch per range."""

    def __init__(self, params, lr=1e-4, beta_1=0.9, beta_2=0.999, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, eps
        self.t = 0
        spans = {}
        for p in params:
            spans.setdefault(id(p.arena), (p.arena, []))[1].append((p.offset, p.offset + (p.size + 3) // 4 * 4))
        self.ranges = []
        for arena, lst in spans.values():
            lst.sort()
            cur_a, cur_b = lst[0]
            for a, b in lst[1:]:
                if a <= cur_b:
                    cur_b = max(cur_b, b)
                else:
                    self.ranges.append((arena, cur_a, cur_b))
                    cur_a, cur_b = a, b
            self.ranges.append((arena, cur_a, cur_b))
        self.m = self.v = None      # allocated on first use (keeps host-only builds possible)

    def _state(self):
        if self.m is None:
            self.m = [ops.zeros(b - a) for _, a, b in self.ranges]
            self.v = [ops.zeros(b - a) for _, a, b in self.ranges]
            self.sched = ops.zeros(4)       # [t, lr_t] kept on the device: the step is CUDA-graph capturable

    def zero_grad(self):
        for arena, a, b in self.ranges:
            ops.zero_(arena.gflat[a:b])

    def grad_buckets(self):
        return [arena.gflat[a:b] for arena, a, b in self.ranges]

    def step(self, grad_scale=1.0):
        self._state()
        self.t += 1
        ops.adam_tick(self.sched, self.lr, self.b1, self.b2)     # t <- t+1; lr_t = lr*sqrt(1-b2^t)/(1-b1^t)
        for (arena, a, b), m, v in zip(self.ranges, self.m, self.v):
            ops.adam_step_dev(arena.flat[a:b], arena.gflat[a:b], m, v, None, self.sched, self.b1, self.b2, self.eps, grad_scale)
            arena.version += 1


# --------------------------------------------------------------------------------------------
# initialisers (Keras 2.1.6 VarianceScaling family), host side
# --------------------------------------------------------------------------------------------
def _fans(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def _truncated_normal(rng, shape, stddev):
    out = rng.normal(0.0, stddev, size=shape)
    bad = np.abs(out) > 2 * stddev
    while bad.any():
        out[bad] = rng.normal(0.0, stddev, size=int(bad.sum()))
        bad = np.abs(out) > 2 * stddev
    return out.astype(np.float32)


def init_weights(rng, shape, kind):
    fan_in, fan_out = _fans(shape)
    if kind == "he_normal":
        return _truncated_normal(rng, shape, math.sqrt(2.0 / fan_in))
    if kind == "glorot_normal":
        return _truncated_normal(rng, shape, math.sqrt(2.0 / (fan_in + fan_out)))
    if kind == "glorot_uniform":
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)
    if kind == "zeros":
        return np.zeros(shape, np.float32)
    if kind == "ones":
        return np.ones(shape, np.float32)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------
class Conv2D:
    """keras Conv2D(filters, k, strides, padding, kernel_initializer) -- NHWC, HWIO kernel."""

    def __init__(self, arena, rng, name, cin, cout, k, stride=1, padding="valid", init="glorot_uniform",
                 use_bias=True):
        self.name, self.cin, self.cout, self.k, self.stride = name, cin, cout, k, stride
        assert padding in ("valid", "same")
        if padding == "same":
            assert stride == 1 and k % 2 == 1
        self.pad = k // 2 if padding == "same" else 0
        self.kernel = arena.add(name + "/kernel", (k, k, cin, cout), init_weights(rng, (k, k, cin, cout), init))
        self.bias = arena.add(name + "/bias", (cout,), np.zeros(cout, np.float32)) if use_bias else None
        self._packed = None          # (arena version, wp_fwd, wp_dgrad)
        self._packed_nc = None
        self._packed_s2d = None
        self._packed_dg = None
        self._folded = None          # ((kernel arena version, BN state version), wp, bias', scale)
        self.bf16_grad = False       # raster-strip path: take / hand on the output gradient in bf16 (FiLM decoder)

    def params(self):
        return [self.kernel] + ([self.bias] if self.bias is not None else [])

    def tc_eligible(self, srcs):
        """128B-swizzled tcgen05 path: channel counts that are multiples of 16 with at least 32 input channels (fewer go
        to the raster-strip kernels); partial 64-channel blocks are zero-padded by TMA and the packed weights.  Stride 1
        (any kernel / padding) or stride 2 with an even kernel and no padding (the discriminator's 4x4 stride-2
        layers).  The first of two concatenated sources must be a whole number of 64-channel blocks."""
        cs = [s.shape[-1] for s in srcs]
        if not (USE_TC and self.k > 1 and self.cout % 16 == 0 and all(c % 16 == 0 for c in cs) and sum(cs) >= 32):
            return False
        if len(cs) > 1 and cs[0] % 64 != 0:
            return False
        return self.stride == 1 or (self.stride == 2 and self.k % 2 == 0 and self.pad == 0)

    def packed(self):
        """bf16 operand copies of the kernel: forward layout + data-gradient layout(s), refreshed IN PLACE after every
        optimizer step that touched the arena (persistent buffers keep a captured CUDA graph valid)"""
        ver = self.kernel.arena.version
        if self._packed is None or self._packed[0] != ver:
            w = self.kernel.data
            old = self._packed
            if self.stride == 1:
                d = ops.pack_conv(w, 1, out=None if old is None else old[2])
            else:       # the four output-parity classes back to back: one data-gradient launch covers them all
                d = ops.pack_conv_s2_all(w, out=None if old is None else old[2])
            self._packed = (ver, ops.pack_conv(w, 0, out=None if old is None else old[1]), d)
        return self._packed[1], self._packed[2]

    def _tc_dgrad(self, g, x_shape, c, off, out_dtype, wp_d):
        """data gradient towards a source with `c` channels starting at channel `off` of the kernel's Cin"""
        N, H, W, _ = x_shape
        k = self.k
        if self.stride == 1:
            out = torch.empty((N, H, W, c), dtype=out_dtype, device=g.device)
            return ops.conv_tc_fwd(g, None, wp_d, None, c, k, k, 1, k - 1 - self.pad, out_dtype, row_off=off, out=out)
        return ops.conv_tc_dgrad_s2(g, wp_d, (N, H, W, c), c, k, k, out_dtype, row_off=off,
                                    rows_per_tap=(self.cin + 63) // 64 * 64)

    def __call__(self, ctx, x, act=None, alpha=0.0, out_dtype=None, bn_stats=False):
        """out_dtype: storage of the convolution output.  The layers that feed a BatchNorm ask for the feature dtype
        (bf16 on the tensor-core path): halves the bytes of the largest tensors of the step; the batch statistics are
        then taken from exactly the values that get normalised."""
        srcs = list(x) if isinstance(x, (list, tuple)) else [x]
        assert sum(s.shape[-1] for s in srcs) == self.cin, (self.name, [tuple(s.shape) for s in srcs], self.cin)
        od = torch.float32 if (out_dtype is None or not USE_TC or not RAW_BF16) else out_dtype
        if self.tc_eligible(srcs) and len(srcs) <= 2:
            return self._call_tc(ctx, srcs, act, alpha, od, bn_stats)
        return self._call_generic(ctx, srcs, act, alpha, od)

    def packed_nc(self):
        """bf16 operands of the narrow-channel tcgen05 kernels (forward + stride-1 data gradient)"""
        ver = self.kernel.arena.version
        if self._packed_nc is None or self._packed_nc[0] != ver:
            old = self._packed_nc
            self._packed_nc = (ver, ops.pack_conv_nc(self.kernel.data, 0, out=None if old is None else old[1]),
                               ops.pack_conv_nc(self.kernel.data, 1, out=None if old is None else old[2]))
        return self._packed_nc[1], self._packed_nc[2]

    def packed_s2d(self):
        """stride-2 layer as a stride-1 layer over 2x2 pixel blocks: rearranged kernel [ceil(k/2)^2 taps, 4*Cin, Cout],
        packed for the raster-strip kernels (narrow layers) or the 128B-swizzled tcgen05 kernels (4*Cin and Cout
        multiples of 64): forward + data-gradient operands, refreshed in place"""
        ver = self.kernel.arena.version
        if self._packed_s2d is None or self._packed_s2d[0] != ver:
            old = self._packed_s2d
            w2 = ops.conv_s2d_weights(self.kernel.data, out=None if old is None else old[3])
            if self._s2d_wide():
                self._packed_s2d = (ver, ops.pack_conv(w2, 0, out=None if old is None else old[1]),
                                    ops.pack_conv(w2, 1, out=None if old is None else old[2]), w2)
            else:
                self._packed_s2d = (ver, ops.pack_conv_nc(w2, 0, out=None if old is None else old[1]),
                                    ops.pack_conv_nc(w2, 1, out=None if old is None else old[2]), w2)
        return self._packed_s2d[1], self._packed_s2d[2]

    def forward_folded(self, srcs, bn, code, out_dtype, narrow=False):
        """predict pass: conv -> BatchNorm(moving statistics) -> [ReLU] as ONE tcgen05 kernel.  In the inference phase
        the normalisation is a per-channel affine map: it is folded into the packed weights (w * gamma * rstd) and the
        bias ((b - mean) * gamma * rstd + beta), the ReLU runs in the epilogue, the output is written once in its final
        dtype.  Refreshed in place whenever the kernel or the moving statistics changed."""
        key = (self.kernel.arena.version, bn.moving_mean.arena.version)
        if self._folded is None or self._folded[0] != key:
            old = self._folded
            scale, bias2 = ops.bn_fold(bn.gamma.data, bn.beta.data, bn.moving_mean.data, bn.moving_var.data,
                                       self.bias.data if self.bias is not None else None, bn.EPS,
                                       out=None if old is None else (old[3], old[2]))
            pack = ops.pack_conv_nc_scaled if narrow else ops.pack_conv_scaled      # raster-strip / swizzled operand
            wp = pack(self.kernel.data, scale, out=None if old is None else old[1])
            self._folded = (key, wp, bias2, scale)
        _, wp, bias2, _ = self._folded
        if narrow:
            return Var(ops.conv_nc_fwd(srcs[0].data, wp, bias2, self.cout, self.k, self.k, self.pad, code, 0.0, out_dtype))
        bs = [s.data if s.data.dtype == torch.bfloat16 else ops.cast(s.data, torch.bfloat16) for s in srcs]
        return Var(ops.conv_tc_fwd(bs[0], bs[1] if len(bs) > 1 else None, wp, bias2, self.cout, self.k, self.k, self.stride,
                                   self.pad, out_dtype, act=code))

    def _s2d_wide(self):
        return (4 * self.cin) % 64 == 0 and self.cout % 64 == 0

    def s2d_eligible(self, srcs):
        """valid stride-2 convolution whose space-to-depth form (4*Cin channels, ceil(k/2) taps, stride 1) runs on
        the tensor cores: conv_tc when 4*Cin and Cout are multiples of 64, conv_nc when the geometry fits"""
        if not (USE_TC and self.stride == 2 and self.pad == 0):
            return False
        H, W = srcs[0].shape[1], srcs[0].shape[2]
        k2 = (self.k + 1) // 2
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        if H2 - k2 + 1 != (H - self.k) // 2 + 1 or W2 - k2 + 1 != (W - self.k) // 2 + 1:
            return False
        if self._s2d_wide():
            return True
        c4 = 4 * self.cin
        return all(ops.nc_supported(c4, self.cout, k2, k2, W2, 0, kind) for kind in (0, 1))

    def _call_s2d(self, ctx, srcs, act, alpha, od=torch.float32):
        code = ACT[act]
        k2 = (self.k + 1) // 2
        wide = self._s2d_wide()
        N, H, W = srcs[0].shape[0], srcs[0].shape[1], srcs[0].shape[2]
        C = sum(s.shape[-1] for s in srcs)
        nc_w = (not wide) and ops.nc_supported(4 * C, self.cout, k2, k2, (W + 1) // 2, 0, 2)
        # Concatenate([a, b]) -> stride-2 convolution (modality encoder): the two sources are rearranged where they lie and
        # the gradient of the rearranged map is split straight back to them (no concatenated copy in either direction)
        fused2 = len(srcs) == 2 and (wide or nc_w) and all(s.grad_dtype == torch.float32 for s in srcs)
        if fused2:
            xin = None
            x2 = ops.space_to_depth2_cat(srcs[0].data, srcs[1].data)
        else:
            xin = srcs[0] if len(srcs) == 1 else concat(ctx, [s if s.data.dtype == torch.float32 else
                                                              _cast_var(ctx, s, torch.float32) for s in srcs])
            x2 = ops.space_to_depth2(xin.data)
        bias = self.bias.data if self.bias is not None else None
        wp_f, wp_d = self.packed_s2d()
        # bf16 storage of the activated output (and a bf16 gradient) only where every consumer below takes it
        lo = od == torch.bfloat16 and code in (ACT_NONE, ACT_RELU, ACT_LRELU) and (wide or nc_w)
        ydt = torch.bfloat16 if lo else torch.float32
        if wide:
            if code in (ACT_NONE, ACT_RELU, ACT_LRELU):
                y = Var(ops.conv_tc_fwd(x2, None, wp_f, bias, self.cout, k2, k2, 1, 0, ydt, act=code, alpha=alpha))
            else:
                raw = ops.conv_tc_fwd(x2, None, wp_f, bias, self.cout, k2, k2, 1, 0, torch.float32)
                y = Var(ops.act_fwd(raw, code, alpha, inplace=True))
        else:
            y = Var(ops.conv_nc_fwd(x2, wp_f, bias, self.cout, k2, k2, 0, code, alpha, ydt))
        if lo:
            y.grad_dtype = torch.bfloat16
        ins = list(srcs) if fused2 else [xin]
        if ctx.rec(*ins, *self.params()):
            y.requires_grad = True

            def bw():
                g = y.grad
                y.grad = None
                if g is None:
                    return
                gb = None
                if lo or (wide and code != ACT_NONE):
                    gb = _act_bwd_to_bf16(g, y.data, code, alpha) if code != ACT_NONE else (
                        g if g.dtype == torch.bfloat16 else ops.cast(g, torch.bfloat16))
                    g = gb if not wide else None         # the raster-strip kernels take the bf16 gradient as it is
                else:
                    if g.dtype != torch.float32:
                        g = ops.cast(g, torch.float32)
                    if code != ACT_NONE:
                        g = ops.act_bwd(g, y.data, code, alpha)
                    gb = ops.cast(g, torch.bfloat16) if wide else None
                if self.kernel.requires_grad:
                    db = self.bias.grad if (self.bias is not None and self.bias.requires_grad) else None
                    if wide:
                        dw2 = ops.zeros(k2, k2, 4 * C, self.cout)
                        ops.conv_tc_wgrad(x2, gb, dw2, 0, k2, k2, 1, 0)
                        ops.conv_s2d_weights_bwd_(self.kernel.grad, dw2)
                        if db is not None:
                            ops.colsum_(gb, db)
                    elif nc_w:
                        dw2 = ops.zeros(k2, k2, 4 * C, self.cout)
                        ops.conv_nc_wgrad(x2, g, dw2, db, 0)
                        ops.conv_s2d_weights_bwd_(self.kernel.grad, dw2)
                    else:
                        xf = xin.data if xin.data.dtype == torch.float32 else ops.cast(xin.data, torch.float32)
                        ops.conv2d_wgrad(xf, g, self.kernel.grad, db, self.stride, self.pad)
                if any(v.requires_grad for v in ins):
                    if wide:
                        dx2 = torch.empty((N, (H + 1) // 2, (W + 1) // 2, 4 * C), dtype=torch.bfloat16, device=gb.device)
                        ops.conv_tc_fwd(gb, None, wp_d, None, 4 * C, k2, k2, 1, k2 - 1, torch.bfloat16, out=dx2)
                    else:
                        dx2 = ops.conv_nc_fwd(g, wp_d, None, 4 * C, k2, k2, k2 - 1, out_dtype=torch.bfloat16)
                    if fused2:
                        ga, gb2 = ops.depth_to_space2_split(dx2, H, W, ins[0].shape[-1], ins[1].shape[-1],
                                                            ins[0].requires_grad, ins[1].requires_grad)
                        accumulate(ins[0], ga)
                        accumulate(ins[1], gb2)
                    else:
                        accumulate(xin, ops.depth_to_space2(dx2, H, W, xin.grad_dtype))

            ctx.tape.record(bw)
        return y

    # ---- narrow layers: raster-strip tcgen05 kernels (stride 1) or the CUDA-core kernels (fp32)
    def _call_generic(self, ctx, srcs, act, alpha, od=torch.float32):
        if self.s2d_eligible(srcs):
            return self._call_s2d(ctx, srcs, act, alpha, od)
        code = ACT[act]
        if (USE_TC and self.k == 1 and self.stride == 1 and len(srcs) == 1 and code == ACT_NONE
                and od == torch.float32 and srcs[0].data.dtype == torch.bfloat16
                and ops.conv1x1_supported(self.cin, self.cout)):
            return self._call_1x1(ctx, srcs[0])
        W = srcs[0].shape[2]
        nc = self.stride == 1 and USE_TC
        nc_f = nc and ops.nc_supported(self.cin, self.cout, self.k, self.k, W, self.pad, 0)
        nc_d = nc and ops.nc_supported(self.cin, self.cout, self.k, self.k, W, self.pad, 1)
        nc_w = nc and ops.nc_supported(self.cin, self.cout, self.k, self.k, W, self.pad, 2)
        # Concatenate([a, b]) -> narrow convolution (locnet, layers/stn_spline.py:104-106): both kernels read the two
        # sources where they lie (a channel group picks its base pointer); only the data gradient is split afterwards
        cat2 = (len(srcs) == 2 and nc_f and nc_w and nc_d and srcs[0].data.dtype == srcs[1].data.dtype
                and srcs[0].shape[-1] % 8 == 0)
        if cat2:
            xin = None
        elif len(srcs) == 1 and nc_f and nc_w:
            xin = srcs[0]                      # f32 or bf16: converted while it is staged
        else:
            f32srcs = [s if s.data.dtype == torch.float32 else _cast_var(ctx, s, torch.float32) for s in srcs]
            xin = f32srcs[0] if len(f32srcs) == 1 else concat(ctx, f32srcs)
        bias = self.bias.data if self.bias is not None else None
        # bf16 output gradients only on request (FiLM decoder under DEC_BF16).  Measured on B200 for the 8 -> 64 / 1 -> 64 first
        # layers (the gradient comes from a BatchNorm backward that could write bf16): the raster-strip kernels stage
        # (pixel, 8-channel) units, so halving the bytes per unit does not make them faster -- weight gradient 191 -> 289 us,
        # 64 -> 8 data gradient 326 -> 596 us -- and fp32 stays
        bf16_bw = self.bf16_grad and nc_f and nc_w and od == torch.bfloat16
        # Wide first layers (8 -> 64 segmentor / discriminator conv1, 1 -> 64 UNet conv1; models/unet.py:95,
        # model_components/segmentor.py:15) whose output is stored in bf16: the BatchNormalization backward writes the
        # output gradient in bf16 when the weight-gradient kernel stages rows with bulk copies (141 us against 157 us from
        # fp32), and the 64 -> 8 data gradient then runs on the swizzled tcgen05 kernel (94 us against 334 us on the
        # raster-strip kernel).  Same numbers either way: the raster-strip kernels round the gradient to bf16 while staging.
        tc_dg = nc and self.k > 1 and self.cout % 16 == 0 and self.cout >= 32 and self.cin % 8 == 0
        wide_bf16 = (WIDE_BF16_GRAD and nc_f and nc_w and not cat2 and od == torch.bfloat16 and self.cout >= 32
                     and self.cout % 16 == 0 and xin.data.dtype == torch.float32 and ctx.training
                     and (tc_dg or not xin.requires_grad)
                     and ops.nc_wgrad_stages_raw(tuple(xin.shape), torch.float32, self.cout, self.k, self.pad))
        bf16_bw = bf16_bw or wide_bf16
        if cat2:
            y = Var(ops.conv_nc_fwd_cat(srcs[0].data, srcs[1].data, self.packed_nc()[0], bias, self.cout, self.k, self.k,
                                        self.pad, code, alpha, od))
        elif nc_f:
            y = Var(ops.conv_nc_fwd(xin.data, self.packed_nc()[0], bias, self.cout, self.k, self.k, self.pad, code, alpha, od),
                    grad_dtype=torch.bfloat16 if bf16_bw else torch.float32)
        else:
            y = Var(ops.conv2d_fwd(xin.data, self.kernel.data, bias, self.stride, self.pad, code, alpha))
        ins = list(srcs) if cat2 else [xin]
        if ctx.rec(*ins, *self.params()):
            y.requires_grad = True
            tape_x = xin

            def bw():
                if (y._grad2 is not None and code != ACT_NONE and y._grad.dtype == y._grad2.dtype == y.data.dtype == torch.float32
                        and y.data.numel() % 4 == 0):
                    # two consumers (FiLM block: conv2 and the residual Add): sum + activation backward in one pass
                    g1, g2, owned = y.take_grad_pair()
                    g = ops.add_act_bwd(g1, g2, y.data, code, alpha, out=g1 if owned else None)
                else:
                    g = y.grad
                    y.grad = None
                    if g is None:
                        return
                    if bf16_bw and g.dtype == torch.bfloat16 and g.numel() % 8 == 0:
                        if code != ACT_NONE:                  # bf16 gradient x bf16 activation output, no fp32 intermediate
                            g = ops.act_bwd_bf16io(g, y.data, code, alpha)
                    else:
                        if g.dtype != torch.float32:
                            g = ops.cast(g, torch.float32)
                        if code != ACT_NONE:
                            g = ops.act_bwd(g, y.data, code, alpha)
                if self.kernel.requires_grad:
                    db = self.bias.grad if (self.bias is not None and self.bias.requires_grad) else None
                    if cat2:
                        ops.conv_nc_wgrad_cat(ins[0].data, ins[1].data, g, self.kernel.grad, db, self.pad)
                    elif nc_w:
                        ops.conv_nc_wgrad(tape_x.data, g, self.kernel.grad, db, self.pad)
                    else:
                        ops.conv2d_wgrad(tape_x.data, g, self.kernel.grad, db, self.stride, self.pad)
                if cat2:
                    if any(v.requires_grad for v in ins):
                        dx = ops.conv_nc_fwd(g, self.packed_nc()[1], None, self.cin, self.k, self.k, self.k - 1 - self.pad)
                        off = 0
                        for v in ins:
                            if v.requires_grad:
                                accumulate(v, ops.slice_channels(dx, off, v.shape[-1]))
                            off += v.shape[-1]
                elif tape_x.requires_grad:
                    if tc_dg and g.dtype == torch.bfloat16:
                        ver = self.kernel.arena.version
                        if self._packed_dg is None or self._packed_dg[0] != ver:
                            self._packed_dg = (ver, ops.pack_conv(self.kernel.data, 1,
                                                                  out=None if self._packed_dg is None else self._packed_dg[1]))
                        dx = ops.conv_tc_fwd(g, None, self._packed_dg[1], None, self.cin, self.k, self.k, 1,
                                             self.k - 1 - self.pad, tape_x.grad_dtype)
                    elif nc_d:
                        dx = ops.conv_nc_fwd(g, self.packed_nc()[1], None, self.cin, self.k, self.k, self.k - 1 - self.pad,
                                             out_dtype=tape_x.grad_dtype)
                    elif tc_dg:
                        # few inputs <- many outputs (SPADE's 8 -> 128 anatomy convolution, layers/spade.py:29): the data
                        # gradient is a wide-in / narrow-out convolution for the swizzled tcgen05 kernel
                        ver = self.kernel.arena.version
                        if self._packed_dg is None or self._packed_dg[0] != ver:
                            self._packed_dg = (ver, ops.pack_conv(self.kernel.data, 1,
                                                                  out=None if self._packed_dg is None else self._packed_dg[1]))
                        dx = ops.conv_tc_fwd(ops.cast(g, torch.bfloat16), None, self._packed_dg[1], None, self.cin, self.k,
                                             self.k, 1, self.k - 1 - self.pad, tape_x.grad_dtype)
                    else:
                        dx = ops.conv2d_dgrad(g, self.kernel.data, tuple(tape_x.shape), self.stride, self.pad)
                    accumulate(tape_x, dx)

            ctx.tape.record(bw)
        return y

    # ---- pointwise heads on a 64-channel bf16 map (anatomy 64 -> 8, segmentor 64 -> 5): HBM-stream kernels
    def _call_1x1(self, ctx, xin):
        bias = self.bias.data if self.bias is not None else None
        y = Var(ops.conv1x1_fwd(xin.data, self.kernel.data, bias), grad_dtype=torch.float32)
        if ctx.rec(xin, *self.params()):
            y.requires_grad = True

            def bw():
                g = y.grad
                y.grad = None
                if g is None:
                    return
                if g.dtype != torch.float32:
                    g = ops.cast(g, torch.float32)
                if self.kernel.requires_grad:
                    db = self.bias.grad if (self.bias is not None and self.bias.requires_grad) else None
                    ops.conv1x1_wgrad(xin.data, g, self.kernel.grad, db)
                if xin.requires_grad:
                    accumulate(xin, ops.conv1x1_dgrad(g, self.kernel.data))

            ctx.tape.record(bw)
        return y

    # ---- tcgen05 path (bf16 operands, fp32 accumulate in TMEM)
    def _call_tc(self, ctx, srcs, act, alpha, od=torch.float32, bn_stats=False):
        bsrcs = [s if s.data.dtype == torch.bfloat16 else _cast_var(ctx, s, torch.bfloat16) for s in srcs]
        wp_f, wp_d = self.packed()
        bias = self.bias.data if self.bias is not None else None
        x0 = bsrcs[0]
        x1 = bsrcs[1] if len(bsrcs) > 1 else None
        code = ACT[act]
        fuse = code in (ACT_RELU, ACT_LRELU)      # the activation runs in the epilogue: y = act(conv + bias), stored once
        acc = None
        if bn_stats and code == ACT_NONE and od == torch.bfloat16 and self.cout >= 128:
            # a training-phase BatchNorm follows: its batch statistics come out of this kernel's epilogue.  Measured on
            # B200 (B = 32): the statistics add 6 - 12 us to layers with >= 128 outputs and replace a 19 - 27 us pass;
            # on the 64-channel 224^2 layers the epilogue warps are the critical path (143 -> 201 us against a 38 us
            # statistics pass), so those keep the separate pass
            raw, acc = ops.conv_tc_fwd_bn(x0.data, None if x1 is None else x1.data, wp_f, bias, self.cout, self.k, self.k,
                                          self.stride, self.pad)
        else:
            raw = ops.conv_tc_fwd(x0.data, None if x1 is None else x1.data, wp_f, bias, self.cout, self.k, self.k,
                                  self.stride, self.pad, od if (fuse or code == ACT_NONE) else torch.float32,
                                  act=code if fuse else ACT_NONE, alpha=alpha)
        # an fp32-stored activated output (last discriminator layer, feeds a Dense) takes its gradient in fp32 and
        # converts while the activation backward runs; everything else hands bf16 to the gradient kernels
        y = Var(raw, grad_dtype=torch.float32 if (fuse and raw.dtype == torch.float32) else torch.bfloat16)
        y.bn_acc = acc
        rec = ctx.rec(*bsrcs, *self.params())
        if rec:
            y.requires_grad = True
            want_db = self.bias is not None and self.bias.requires_grad
            if want_db and not fuse:
                y.bias_sink = self.bias       # a following BatchNorm backward folds the bias gradient in

            def bw():
                g = y.grad
                y.grad = None
                if g is None:
                    return
                if fuse:
                    g = _act_bwd_to_bf16(g, y.data, code, alpha)
                    if want_db:
                        ops.colsum_(g, self.bias.grad)
                else:
                    if g.dtype != torch.bfloat16:
                        g = ops.cast(g, torch.bfloat16)
                    if y.bias_sink is not None:   # nobody consumed it: reduce here
                        ops.colsum_(g, self.bias.grad)
                off = 0
                for s in bsrcs:
                    c = s.shape[-1]
                    if self.kernel.requires_grad:
                        ops.conv_tc_wgrad(s.data, g, self.kernel.grad, off, self.k, self.k, self.stride, self.pad)
                    if s.requires_grad:
                        accumulate(s, self._tc_dgrad(g, tuple(s.shape), c, off, s.grad_dtype, wp_d))
                    off += c

            ctx.tape.record(bw)
        if code != ACT_NONE and not fuse:
            return activation(ctx, y, act, alpha)
        return y


def _act_bwd_to_bf16(g, y, code, alpha):
    """bf16(g * act'(y)) for an activation output y stored in bf16 or fp32 (one pass whenever the dtypes pair up)"""
    if g.dtype == torch.bfloat16 and y.dtype == torch.bfloat16 and g.numel() % 8 == 0:
        return ops.act_bwd_bf16io(g, y, code, alpha)
    if g.dtype == torch.float32 and y.dtype == torch.float32 and g.numel() % 4 == 0:
        return ops.act_bwd_bf16(g, y, code, alpha)
    gf = g if g.dtype == torch.float32 else ops.cast(g, torch.float32)
    yf = y if y.dtype == torch.float32 else ops.cast(y, torch.float32)
    return ops.cast(ops.act_bwd(gf, yf, code, alpha), torch.bfloat16)


def _cast_var(ctx, v, dtype):
    out = Var(ops.cast(v.data, dtype))
    if ctx.rec(v):
        out.requires_grad = True

        def bw():
            g = out.grad
            out.grad = None
            if g is not None:
                accumulate(v, g)

        ctx.tape.record(bw)
    return out


class BatchNorm:
    """keras BatchNormalization() with its defaults (axis -1, momentum .99, epsilon 1e-3)."""
    EPS = 1e-3
    MOMENTUM = 0.99

    def __init__(self, arena, state, name, c):
        self.name, self.c = name, c
        self.gamma = arena.add(name + "/gamma", (c,), np.ones(c, np.float32))
        self.beta = arena.add(name + "/beta", (c,), np.zeros(c, np.float32))
        self.moving_mean = state.add(name + "/moving_mean", (c,), np.zeros(c, np.float32))
        self.moving_var = state.add(name + "/moving_variance", (c,), np.ones(c, np.float32))

    def params(self):
        return [self.gamma, self.beta]

    def __call__(self, ctx, x, act=None, out_dtype=torch.float32):
        code = ACT[act]
        if ctx.training:
            if x.bn_acc is not None:       # the producing convolution already summed its stored outputs
                mean, rstd = ops.bn_finalize_acc(x.bn_acc, x.data.numel() // self.c, self.EPS, self.MOMENTUM,
                                                 self.moving_mean.data, self.moving_var.data)
                x.bn_acc = None
            else:
                mean, rstd = ops.bn_stats_finalize(x.data, self.EPS, self.MOMENTUM, self.moving_mean.data,
                                                   self.moving_var.data)
            self.moving_mean.arena.version += 1        # folded inference copies (Conv2D.forward_folded) are stale now
        else:
            mean, rstd = self.moving_mean.data, ops.bn_rstd_from_var(self.moving_var.data, self.EPS)
        y = Var(ops.bn_apply(x.data, mean, rstd, self.gamma.data, self.beta.data, code, out_dtype))
        if ctx.rec(x, *self.params()):
            assert ctx.training, "BatchNorm backward is only recorded in training mode"
            y.requires_grad = True

            def bw():
                g = y.grad
                y.grad = None
                if g is None:
                    return
                tr = self.gamma.requires_grad
                sink = x.bias_sink
                x.bias_sink = None
                dx = ops.bn_bwd(g, x.data, mean, rstd, self.gamma.data, self.beta.data, code,
                                self.gamma.grad if tr else None, self.beta.grad if tr else None,
                                dx_dtype=x.grad_dtype, dbias_prev=None if sink is None else sink.grad)
                accumulate(x, dx)

            ctx.tape.record(bw)
        return y


class InstanceNorm:
    """keras_contrib InstanceNormalization() with its defaults (axis=None, epsilon 1e-3, center and scale with gamma,
    beta of shape (1,)): utils/model_utils.py:6-12 normalise('instance').  Per-sample statistics over H, W, C jointly,
    the same in the training and the inference phase (no moving statistics)."""
    EPS = 1e-3

    def __init__(self, arena, name):
        self.name = name
        self.gamma = arena.add(name + "/gamma", (1,), np.ones(1, np.float32))
        self.beta = arena.add(name + "/beta", (1,), np.zeros(1, np.float32))

    def params(self):
        return [self.gamma, self.beta]

    def __call__(self, ctx, x, act=None, out_dtype=torch.float32):
        code = ACT[act]
        xin = x if x.data.dtype == torch.float32 else _cast_var(ctx, x, torch.float32)
        acc = ops.in_stats(xin.data)
        out = ops.in_affine_fwd(xin.data, acc, self.gamma.data, self.beta.data, code, 0.0, self.EPS)
        y = Var(out)
        if ctx.rec(xin, *self.params()):
            y.requires_grad = True

            def bw():
                g = y.grad
                y.grad = None
                if g is None:
                    return
                if g.dtype != torch.float32:
                    g = ops.cast(g, torch.float32)
                tr = self.gamma.requires_grad
                # x.bias_sink stays with the producing convolution: under joint H,W,C statistics its bias does not cancel
                dx = ops.in_affine_bwd(g, xin.data, acc, self.gamma.data, self.beta.data,
                                       self.gamma.grad if tr else None, self.beta.grad if tr else None, code, 0.0, self.EPS)
                accumulate(xin, dx)

            ctx.tape.record(bw)
        if out_dtype != torch.float32:
            return _cast_var(ctx, y, out_dtype)
        return y


class Dense:
    def __init__(self, arena, rng, name, cin, cout, init="glorot_uniform", bias_init="zeros"):
        self.name, self.cin, self.cout = name, cin, cout
        self.kernel = arena.add(name + "/kernel", (cin, cout), init_weights(rng, (cin, cout), init))
        self.bias = arena.add(name + "/bias", (cout,), np.zeros(cout, np.float32))

    def params(self):
        return [self.kernel, self.bias]

    def __call__(self, ctx, x, act=None, alpha=0.0):
        x2 = x.data.reshape(x.shape[0], -1)
        assert x2.shape[1] == self.cin, (self.name, tuple(x.shape), self.cin)
        y = Var(ops.dense_fwd(x2, self.kernel.data, self.bias.data))
        if ctx.rec(x, *self.params()):
            y.requires_grad = True

            def bw():
                g = y.grad
                y.grad = None
                if g is None:
                    return
                if self.kernel.requires_grad:
                    ops.dense_bwd_weight(x2, g, self.kernel.grad, self.bias.grad)
                if x.requires_grad:
                    accumulate(x, ops.dense_bwd_data(g, self.kernel.data).view(x.shape))

            ctx.tape.record(bw)
        if ACT[act] != ACT_NONE:
            return activation(ctx, y, act, alpha)
        return y


WIDE_BF16_GRAD = os.environ.get("DAFK_WIDE_BF16", "1") != "0"     # bf16 output gradients into the wide first layers (Conv2D.__call__)
FOLD_BN = True      # predict passes: fold BatchNorm (+ReLU) into the tensor-core convolution that feeds it


def conv_bn(ctx, conv, bn, x, act=None, out_dtype=torch.float32):
    """Conv2D -> BatchNormalization -> activation (models/unet.py:94-101, utils/model_utils.py:15-22,
    model_components/segmentor.py:15-21).  Training phase / taped graphs: three kernels families (convolution,
    statistics, apply).  Inference phase on the tensor-core path: one kernel (Conv2D.forward_folded)."""
    srcs = list(x) if isinstance(x, (list, tuple)) else [x]
    code = ACT[act]
    if isinstance(bn, InstanceNorm):      # no moving statistics: nothing to fold
        return bn(ctx, conv(ctx, x), act, out_dtype)
    if (FOLD_BN and USE_TC and not ctx.training and ctx.tape is None and code in (ACT_NONE, ACT_RELU)
            and len(srcs) <= 2 and conv.stride == 1 and conv.tc_eligible(srcs)):
        return conv.forward_folded(srcs, bn, code, out_dtype)
    if (FOLD_BN and USE_TC and not ctx.training and ctx.tape is None and code in (ACT_NONE, ACT_RELU) and len(srcs) == 1
            and conv.stride == 1 and not conv.tc_eligible(srcs)
            and ops.nc_supported(conv.cin, conv.cout, conv.k, conv.k, srcs[0].shape[2], conv.pad, 0)):
        return conv.forward_folded(srcs, bn, code, out_dtype, narrow=True)      # first layers: 1 -> 64, 8 -> 64
    return bn(ctx, conv(ctx, x, out_dtype=feat_dtype(), bn_stats=ctx.training and ops.FUSE_BN_STATS), act, out_dtype)


# --------------------------------------------------------------------------------------------
# functional ops
# --------------------------------------------------------------------------------------------
def activation(ctx, x, act, alpha=0.0):
    code = ACT[act]
    y = Var(ops.act_fwd(x.data, code, alpha))
    if ctx.rec(x):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            if (x.grad_dtype == torch.bfloat16 and g.dtype == torch.float32 and y.data.dtype == torch.float32
                    and g.numel() % 4 == 0 and x.requires_grad):
                accumulate(x, ops.act_bwd_bf16(g, y.data, code, alpha))      # consumer wants bf16: no fp32 intermediate
            else:
                accumulate(x, ops.act_bwd(g, y.data, code, alpha))

        ctx.tape.record(bw)
    return y


def add(ctx, a, b):
    y = Var(ops.add(a.data, b.data))
    if ctx.rec(a, b):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is not None:
                accumulate(a, g, owned=False)
                accumulate(b, g, owned=False)

        ctx.tape.record(bw)
    return y


def concat(ctx, vs):
    y = Var(ops.concat_channels([v.data for v in vs]))
    if ctx.rec(*vs):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            off = 0
            for v in vs:
                c = v.shape[-1]
                if v.requires_grad:
                    accumulate(v, ops.slice_channels(g, off, c))
                off += c

        ctx.tape.record(bw)
    return y


def concat_rows(ctx, vs):
    """concatenate along the batch axis (one call of a weight-sharing component instead of k calls)"""
    n = sum(v.shape[0] for v in vs)
    out = torch.empty((n,) + tuple(vs[0].shape[1:]), dtype=vs[0].data.dtype, device=vs[0].data.device)
    off = 0
    for v in vs:
        ops.copy_(out[off:off + v.shape[0]], v.data)
        off += v.shape[0]
    y = Var(out)
    if ctx.rec(*vs):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            off = 0
            for v in vs:
                b = v.shape[0]
                if v.requires_grad:
                    accumulate(v, g[off:off + b], owned=False)
                off += b

        ctx.tape.record(bw)
    return y


def split_rows(ctx, x, sizes):
    """views of consecutive batch slices; the backward node gathers the slices' gradients into one tensor"""
    parts, off = [], 0
    for b in sizes:
        parts.append(Var(x.data[off:off + b]))
        off += b
    if ctx.rec(x):
        for p_ in parts:
            p_.requires_grad = True
            p_.grad_dtype = x.grad_dtype

        def bw():
            if all(p_.grad is None for p_ in parts):
                return
            g = torch.empty(x.data.shape, dtype=x.grad_dtype, device=x.data.device)
            off = 0
            for p_, b in zip(parts, sizes):
                if p_.grad is None:
                    ops.zero_(g[off:off + b])
                else:
                    ops.copy_(g[off:off + b], p_.grad)
                    p_.grad = None
                off += b
            accumulate(x, g)

        ctx.tape.record(bw)
    return parts


def slice_channels(ctx, x, off, c):
    y = Var(ops.slice_channels(x.data, off, c))
    if ctx.rec(x):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            full = ops.zeros(*x.shape)
            ops.copy_channels(g, 0, full, off, c)
            accumulate(x, full)

        ctx.tape.record(bw)
    return y


def maxpool2(ctx, x):
    y = Var(ops.maxpool2_fwd(x.data))
    if ctx.rec(x):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is not None:
                if g.dtype != x.data.dtype:
                    g = ops.cast(g, x.data.dtype)
                accumulate(x, ops.maxpool2_bwd(x.data, g))

        ctx.tape.record(bw)
    return y


def upsample2(ctx, x):
    y = Var(ops.upsample2_fwd(x.data))
    if ctx.rec(x):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is not None:
                accumulate(x, ops.upsample2_bwd(g))

        ctx.tape.record(bw)
    return y


def softmax(ctx, x, rounding=False):
    """softmax over the last axis, optionally followed by Rounding (straight-through gradient)."""
    p, r = ops.softmax_fwd(x.data, want_round=rounding)
    y = Var(r if rounding else p)
    if ctx.rec(x):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is not None:
                accumulate(x, ops.softmax_bwd(p, g))     # Rounding gradient is the identity

        ctx.tape.record(bw)
    return y


def rounding(ctx, x):
    y = Var(ops.round_fwd(x.data))
    if ctx.rec(x):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is not None:
                accumulate(x, g, owned=False)      # straight-through: the same buffer flows on

        ctx.tape.record(bw)
    return y


def film(ctx, x, gamma, beta):
    y = Var(ops.film_fwd(x.data, gamma.data, beta.data))
    if ctx.rec(x, gamma, beta):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            dx, dg, db = ops.film_bwd(g, x.data, gamma.data)
            accumulate(x, dx)
            accumulate(gamma, dg)
            accumulate(beta, db)

        ctx.tape.record(bw)
    return y


def film_act_add(ctx, x, gamma, beta, res, act, alpha=0.0):
    """res + act(FiLM(x, gamma, beta)): the three element-wise layers that close a FiLM block
    (model_components/decoder.py:50-54) as one forward and one backward pass"""
    code = ACT[act]
    y = Var(ops.film_act_add_fwd(x.data, gamma.data, beta.data, res.data, code, alpha))
    if ctx.rec(x, gamma, beta, res):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            if x.requires_grad or gamma.requires_grad or beta.requires_grad:
                dx, dg, db = ops.film_act_add_bwd(g, x.data, gamma.data, beta.data, code, alpha)
                accumulate(x, dx)
                accumulate(gamma, dg)
                accumulate(beta, db)
            accumulate(res, g, owned=False)

        ctx.tape.record(bw)
    return y


def maximum(ctx, a, b):
    y = Var(ops.max_fwd(a.data, b.data))
    if ctx.rec(a, b):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            da, db = ops.max_bwd(a.data, b.data, g)
            accumulate(a, da)
            accumulate(b, db)

        ctx.tape.record(bw)
    return y


def tps_warp(ctx, vol, theta, cp=(5, 5)):
    out, _ = ops.tps_warp_fwd(vol.data, theta.data, cp)
    y = Var(out)
    if ctx.rec(vol, theta):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            dvol, dtheta = ops.tps_warp_bwd(vol.data, theta.data, g, cp, need_dvol=vol.requires_grad)
            if vol.requires_grad:
                accumulate(vol, dvol)
            accumulate(theta, dtheta)

        ctx.tape.record(bw)
    return y


def reshape(ctx, x, shape):
    y = Var(x.data.view(shape))
    if ctx.rec(x):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is not None:
                accumulate(x, g.view(x.shape), owned=False)

        ctx.tape.record(bw)
    return y


def resize_nn(ctx, x, ho, wo):
    y = Var(ops.resize_nn_fwd(x.data, ho, wo))
    if ctx.rec(x):
        y.requires_grad = True
        H, W = x.shape[1], x.shape[2]

        def bw():
            g = y.grad
            y.grad = None
            if g is not None:
                accumulate(x, ops.resize_nn_bwd(g, H, W))

        ctx.tape.record(bw)
    return y


def spade_norm(ctx, x, gamma, beta, act="lrelu", alpha=0.2):
    """InstanceNormalization(axis=None, no affine) -> SPADE_COND -> activation, fused (layers/spade.py:26-32,52-55)."""
    code = ACT[act]
    acc = ops.in_stats(x.data)
    y = Var(ops.spade_fwd(x.data, acc, gamma.data, beta.data, code, alpha))
    if ctx.rec(x, gamma, beta):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            dx, dg, db = ops.spade_bwd(g, x.data, acc, gamma.data, beta.data, code, alpha)
            accumulate(x, dx)
            accumulate(gamma, dg)
            accumulate(beta, db)

        ctx.tape.record(bw)
    return y


def spade_cond(ctx, x, gamma, beta):
    """SPADE_COND()([x, gamma, beta]) = x*(1+gamma)+beta on an already-normalised x (layers/spade.py:41-58)"""
    y = Var(ops.spade_cond_fwd(x.data, gamma.data, beta.data))
    if ctx.rec(x, gamma, beta):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            dx, dg = ops.spade_cond_bwd(g, x.data, gamma.data)
            accumulate(x, dx)
            accumulate(gamma, dg)
            accumulate(beta, g, owned=False)

        ctx.tape.record(bw)
    return y


# --------------------------------------------------------------------------------------------
# losses: each adds weight*value to ``loss_buf[slot]`` and seeds the gradient of its input
# --------------------------------------------------------------------------------------------
def loss_seg(ctx, pred, target, nch, use_bce, weight, loss_slot):
    g = ops.segloss(pred.data, target, nch, use_bce, weight, loss_slot, want_grad=ctx.rec(pred))
    if g is not None:
        accumulate(pred, g)


def loss_l1l2(ctx, pred, target, kind, weight, loss_slot, cval=0.0):
    g = ops.l1l2_loss(pred.data, target, kind, weight, loss_slot, cval=cval, want_grad=ctx.rec(pred))
    if g is not None:
        accumulate(pred, g)


def pair_dice(ctx, a, b):
    """model_components/balancer.py:33-38: per-sample Dice overlap [B,1], differentiable in both anatomies"""
    out, ws = ops.pair_dice(a.data, b.data, want_ws=True)
    y = Var(out)
    if ctx.rec(a, b):
        y.requires_grad = True

        def bw():
            g = y.grad
            y.grad = None
            if g is None:
                return
            da, db = ops.pair_dice_bwd(a.data, b.data, ws, g, a.requires_grad, b.requires_grad)
            if da is not None:
                accumulate(a, da)
            if db is not None:
                accumulate(b, db)

        ctx.tape.record(bw)
    return y


def loss_pairs(ctx, kind, preds, target, weights, weight, loss_slot, nch=None):
    """Add()([Multiply()([w_j, PerSampleLoss([target, pred_j])]) ...]) with loss costs.ypred (models/dafnet.py:283-315):
    loss += weight * mean_b sum_j w[b,j] * L_j[b].  kind 'seg' = make_combined_dice_bce_perbatch, 'mae' =
    mae_single_input.  `weights` is the Balancer output Var [B,P].  Terminal node: the gradients towards the
    predictions and the weights are produced here."""
    P, B = len(preds), preds[0].shape[0]
    L = torch.empty((P, B), dtype=torch.float32, device=preds[0].data.device)
    wss = []
    for j, p in enumerate(preds):
        pd = p.data if p.data.dtype == torch.float32 else ops.cast(p.data, torch.float32)
        wss.append((pd, ops.segloss_pb_fwd(pd, target, nch, L[j]) if kind == "seg" else ops.mae_pb_fwd(pd, target, L[j])))
    need = ctx.rec(*preds) or ctx.rec(weights)
    dw, coef = ops.pair_combine(weights.data, L, weight, loss_slot, want_dw=need)
    if not need:
        return
    if weights.requires_grad:
        accumulate(weights, dw)
    for j, p in enumerate(preds):
        if p.requires_grad:
            pd, ws = wss[j]
            g = ops.segloss_pb_bwd(pd, target, nch, ws, coef[j]) if kind == "seg" else ops.mae_pb_bwd(pd, target, coef[j])
            accumulate(p, g)


def vae_sample(ctx, mu, logvar, eps, kl_weight, loss_slot):
    """z = mu + exp(.5 lv) eps (utils/sdnet_utils.py:9-21) and the KL output whose mean is the
    `Enc_Modality` loss (costs.py:186-195); the KL gradient is folded into the backward of z."""
    z, klv = ops.vae_fwd(mu.data, logvar.data, eps, kl_weight, loss_slot)
    zv = Var(z)
    if ctx.rec(mu, logvar):
        zv.requires_grad = True

        def bw():
            g = zv.grad
            zv.grad = None
            dmu, dlv = ops.vae_bwd(mu.data, logvar.data, eps, g, kl_weight)
            accumulate(mu, dmu)
            accumulate(logvar, dlv)

        ctx.tape.record(bw)
    return zv, klv

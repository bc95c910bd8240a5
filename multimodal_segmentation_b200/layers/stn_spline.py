"""Thin-plate-spline spatial transformer (reference: layers/stn_spline.py:14-120).

ThinPlateSpline2D(input_volume_shape, cp_dims, num_channels, inverse=False, order=2)([vol, cp_offsets])
The reference fits one spline per sample inside tf.map_fn (sequential 28x28 LU + [H*W x 25] matmul)
and then calls tf.contrib.resampler.  With inverse=False (the only mode used, anatomy_fuser.py:30)
the system matrix depends only on the constant 5x5 grid, so the fit collapses to a constant
25x25 / 3x25 matrix applied to the offsets; evaluation and the bilinear gather are ONE fused kernel.
inverse=True falls back to the general batched solve + apply + resampler kernels (forward only).
"""
import numpy as np
import torch

from .. import engine as E
from .. import ops
from ..keras_like import BuildScope, Model
from .interpolate_spline import interpolate_spline


def nDgrid(dims, normalise=True, center=False, dtype="float32"):
    """stn_spline.py:70-91 -> numpy [1, prod(dims), len(dims)]"""
    if len(dims) == 2:
        grid = np.expand_dims(np.mgrid[:dims[0], :dims[1]].reshape((2, -1)).T, 0)
    elif len(dims) == 3:
        grid = np.expand_dims(np.mgrid[:dims[0], :dims[1], :dims[2]].reshape((3, -1)).T, 0)
    else:
        raise ValueError(dims)
    if normalise:
        grid = grid / (1. * (np.array([[dims]]) - 1))
        if center:
            grid = (grid - 1) * 2
    return grid.astype(dtype)


class ThinPlateSpline2D(object):
    def __init__(self, input_volume_shape, cp_dims, num_channels, inverse=False, order=2, **kwargs):
        self.vol_shape = tuple(input_volume_shape)
        self.cp_dims = tuple(cp_dims)
        self.num_channels = num_channels
        self.inverse = inverse
        self.order = order
        self.name = kwargs.get("name", "thin_plate_spline2d")

    def __call__(self, ctx, args):
        vol, cp_offsets = args
        if not self.inverse and self.order == 2:
            return E.tps_warp(ctx, vol, cp_offsets, self.cp_dims)
        if ctx.rec(vol, cp_offsets):
            raise NotImplementedError("inverse/other-order TPS is forward-only (not used by the reference models)")
        B = vol.shape[0]
        H, W = self.vol_shape
        cp = torch.from_numpy(nDgrid(self.cp_dims)).cuda().expand(B, -1, -1).contiguous()
        q = torch.from_numpy(nDgrid(self.vol_shape)).cuda()
        warped = ops.add(cp, cp_offsets.data.contiguous())
        if self.inverse:
            locs = interpolate_spline(warped, cp, q.expand(B, -1, -1).contiguous(), self.order)
        else:
            locs = interpolate_spline(cp, warped, q.expand(B, -1, -1).contiguous(), self.order)
        # reverse (row,col)->(x,y), scale by [W-1, H-1] (stn_spline.py:61-64): done on the host-visible view
        xy = torch.stack([locs[..., 1] * (W - 1), locs[..., 0] * (H - 1)], -1).contiguous()
        out = ops.resampler_fwd(vol.data, xy)
        return E.Var(out.view(B, H, W, self.num_channels))


def build_locnet(input_shape1, input_shape2, output_shape):
    """stn_spline.py:94-120: concat -> [conv5x5(20) + LeakyReLU(.3) + maxpool] x2 -> conv5x5(20) + LeakyReLU
    -> Flatten -> Dense(100, tanh) -> Dense(output_shape, zeros) -> Reshape(output_shape/2, 2)"""
    scope = BuildScope.current()
    a, r = scope.arena, scope.rng
    cin = input_shape1[-1] + input_shape2[-1]
    c1 = E.Conv2D(a, r, "loc_conv1", cin, 20, 5, 1, "valid")
    c2 = E.Conv2D(a, r, "loc_conv2", 20, 20, 5, 1, "valid")
    c3 = E.Conv2D(a, r, "loc_conv3", 20, 20, 5, 1, "valid")

    def hw(n):
        n = (n - 4) // 2
        n = (n - 4) // 2
        return n - 4
    flat = hw(input_shape1[0]) * hw(input_shape1[1]) * 20
    d1 = E.Dense(a, r, "loc_dense1", flat, 100)
    d2 = E.Dense(a, r, "loc_theta", 100, output_shape, "zeros")

    def fwd(ctx, x1, x2):
        # Concatenate([x1, x2]): the first convolution reads the two maps where they lie
        l = E.maxpool2(ctx, c1(ctx, [x1, x2], "lrelu", 0.3))
        l = E.maxpool2(ctx, c2(ctx, l, "lrelu", 0.3))
        l = c3(ctx, l, "lrelu", 0.3)
        l = d1(ctx, l, "tanh")
        theta = d2(ctx, l)
        return E.reshape(ctx, theta, (theta.shape[0], output_shape // 2, 2))

    return Model("stn_locnet", [c1, c2, c3, d1, d2], fwd, [tuple(input_shape1), tuple(input_shape2)],
                 [(output_shape // 2, 2)], scope)

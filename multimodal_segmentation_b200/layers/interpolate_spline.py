"""Polyharmonic spline interpolation (reference: layers/interpolate_spline.py:30-278, itself a copy of
tf.contrib.image.interpolate_spline).  Device tensors in, device tensors out.

    interpolate_spline(train_points[b,n,d], train_values[b,n,k], query_points[b,m,d], order,
                       regularization_weight=0.0) -> [b,m,k]

solve: one warp per (n+d+1)^2 system, LU with partial pivoting in shared memory (tf.matrix_solve);
apply: one thread per query point.  d = 2, n + 3 <= 32, k <= 8.
"""
from .. import ops

EPSILON = 0.0000000001


def _solve_interpolation(train_points, train_values, order, regularization_weight):
    return ops.tps_solve(train_points.contiguous(), train_values.contiguous(), order, regularization_weight)


def _apply_interpolation(query_points, train_points, w, v, order):
    return ops.tps_apply(query_points.contiguous(), train_points.contiguous(), w, v, order)


def interpolate_spline(train_points, train_values, query_points, order, regularization_weight=0.0,
                       name="interpolate_spline"):
    if train_points.shape[-1] != 2:
        raise ValueError("only d == 2 is supported by the CUDA path")
    B = max(train_points.shape[0], train_values.shape[0])
    if train_points.shape[0] != B:
        train_points = train_points.expand(B, -1, -1).contiguous()
    if train_values.shape[0] != B:
        train_values = train_values.expand(B, -1, -1).contiguous()
    w, v = _solve_interpolation(train_points, train_values, order, regularization_weight)
    return _apply_interpolation(query_points, train_points, w, v, order)

"""Balancer (reference: model_components/balancer.py:11-38): Dice(x1, x_i) for i=2..4 -> Dense(5, relu)
-> Dense(n_pairs, name='beta') -> softmax.  Used by the automated-pairing trainers (models/dafnet.py:283-287,352-361);
the overlap is differentiable in both anatomies (engine.pair_dice), so the pairing weights train the encoders too."""
from .. import engine as E
from .. import ops
from ..keras_like import BuildScope, Model


def dice(y):
    """balancer.py:33-38 on device tensors -> [B,1]"""
    y_true, y_pred = y
    return ops.pair_dice(y_true, y_pred)


def build(conf):
    scope = BuildScope.current()
    a, r = scope.arena, scope.rng
    d1 = E.Dense(a, r, "bal_dense", 3, 5)
    beta = E.Dense(a, r, "beta", 5, conf.n_pairs)

    def fwd(ctx, x1, x2, x3, x4):
        overlap = [E.pair_dice(ctx, x1, x) for x in (x2, x3, x4)]
        l = E.concat(ctx, overlap)
        l = d1(ctx, l, "relu")
        w = beta(ctx, l)
        return E.softmax(ctx, w)

    shp = tuple(conf.anatomy_encoder.output_shape)
    return Model("Balancer", [d1, beta], fwd, [shp] * 4, [(conf.n_pairs,)], scope)

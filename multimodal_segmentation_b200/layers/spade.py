"""SPADE conditioning (reference: layers/spade.py:7-58).

spade_block(conf, anatomy, layer, fin, fout):
    l = conv(fmiddle)(lrelu(.2)(_spade(layer, fin))); l = conv(fout)(lrelu(.2)(_spade(l, fmiddle)))
    shortcut = conv1x1(fout, no bias)(_spade(layer, fin)) if fin != fout else layer ; return shortcut + l
_spade(layer, f): InstanceNormalization(axis=None, no affine)(layer); anatomy resized (nearest) to the
    layer resolution -> conv3x3(128, relu) -> gamma, beta = conv3x3(f) ; SPADE_COND: x*(1+gamma)+beta

The per-sample normalisation, the modulation and the following LeakyReLU(0.2) run as ONE fused
bandwidth kernel (statistics pass + apply pass).
"""
from .. import engine as E


class SPADE_COND(object):
    """SPADE_COND()([x, gamma, beta]) = x * (1 + gamma) + beta on an already-normalised input (spade.py:41-58), no
    activation: one kernel each way (``dafk_spade_cond_fwd/bwd``).  The decoder uses the fused ``_Spade`` below, which
    also carries the instance normalisation and the LeakyReLU."""

    def __init__(self, **kwargs):
        self.name = kwargs.get("name", "spade_cond")

    def params(self):
        return []

    def __call__(self, ctx, x=None):
        if x is None:                      # keras call style: SPADE_COND()([x, gamma, beta]) -> no-gradient context
            ctx, x = E.Ctx(None, training=False), ctx
        h, gamma, beta = [_as_f32(ctx, v if isinstance(v, E.Var) else E.Var(v)) for v in x]
        return E.spade_cond(ctx, h, gamma, beta)

    def compute_output_shape(self, input_shape):
        return input_shape[0] if isinstance(input_shape, (list, tuple)) and isinstance(input_shape[0], (list, tuple)) \
            else input_shape


class _Spade:
    def __init__(self, scope, name, ca, f):
        a, r = scope.arena, scope.rng
        self.shared = E.Conv2D(a, r, name + "_shared", ca, 128, 3, 1, "same")
        self.gamma = E.Conv2D(a, r, name + "_gamma", 128, f, 3, 1, "same")
        self.beta = E.Conv2D(a, r, name + "_beta", 128, f, 3, 1, "same")

    def layers(self):
        return [self.shared, self.gamma, self.beta]

    def __call__(self, ctx, anatomy, layer, act):
        h, w = layer.shape[1], layer.shape[2]
        an = anatomy if (anatomy.shape[1] == h and anatomy.shape[2] == w) else E.resize_nn(ctx, anatomy, h, w)
        an = self.shared(ctx, an, "relu")
        g = _as_f32(ctx, self.gamma(ctx, an))
        b = _as_f32(ctx, self.beta(ctx, an))
        return E.spade_norm(ctx, _as_f32(ctx, layer), g, b, act, 0.2)


def _as_f32(ctx, v):
    import torch
    return v if v.data.dtype == torch.float32 else E._cast_var(ctx, v, torch.float32)


class SpadeBlock:
    def __init__(self, scope, name, ca, fin, fout):
        a, r = scope.arena, scope.rng
        fmid = min(fin, fout)
        self.learn_shortcut = fin != fout
        self.s1 = _Spade(scope, name + "_s1", ca, fin)
        self.c1 = E.Conv2D(a, r, name + "_conv1", fin, fmid, 3, 1, "same")
        self.s2 = _Spade(scope, name + "_s2", ca, fmid)
        self.c2 = E.Conv2D(a, r, name + "_conv2", fmid, fout, 3, 1, "same")
        if self.learn_shortcut:
            self.ss = _Spade(scope, name + "_ss", ca, fin)
            self.cs = E.Conv2D(a, r, name + "_convs", fin, fout, 1, 1, "same", use_bias=False)

    def layers(self):
        out = self.s1.layers() + [self.c1] + self.s2.layers() + [self.c2]
        if self.learn_shortcut:
            out += self.ss.layers() + [self.cs]
        return out

    def __call__(self, ctx, anatomy, layer):
        l = self.s1(ctx, anatomy, layer, "lrelu")
        l = self.c1(ctx, l)
        l = self.s2(ctx, anatomy, l, "lrelu")
        l = self.c2(ctx, l)
        sc = layer
        if self.learn_shortcut:
            sc = self.cs(ctx, self.ss(ctx, anatomy, layer, None))
        return E.add(ctx, _as_f32(ctx, sc), _as_f32(ctx, l))


def spade_block(conf, anatomy_input, layer, fin, fout):
    """functional form with the reference signature; builds a fresh block in the current scope"""
    from ..keras_like import BuildScope
    blk = SpadeBlock(BuildScope.current(), "spade_block", anatomy_input.shape[-1], fin, fout)
    return blk

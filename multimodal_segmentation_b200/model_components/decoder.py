"""Decoder (reference: model_components/decoder.py:12-81): FiLM or SPADE conditioning of the anatomy
on the modality factor z, followed by Conv2D(1, 1, tanh, glorot_normal)."""
import torch

from .. import engine as E
from ..keras_like import BuildScope, Model
from ..layers.film import FiLM
from ..layers.spade import SpadeBlock


class _FilmLayer:
    """decoder.py:44-54: l1 = lrelu(conv(x)); l2 = lrelu(FiLM(conv(l1), gamma(z), beta(z))); out = l1 + l2"""

    def __init__(self, scope, name, num_z):
        a, r = scope.arena, scope.rng
        self.c1 = E.Conv2D(a, r, name + "_conv1", 8, 8, 3, 1, "same")
        self.c2 = E.Conv2D(a, r, name + "_conv2", 8, 8, 3, 1, "same")
        self.g = E.Dense(a, r, name + "_gamma", num_z, 8)
        self.b = E.Dense(a, r, name + "_beta", num_z, 8)
        self.film = FiLM()
        self.c1.bf16_grad = self.c2.bf16_grad = True      # only takes effect with engine.DEC_BF16

    def layers(self):
        return [self.c1, self.c2, self.g, self.b]

    def __call__(self, ctx, x, z):
        l1 = self.c1(ctx, x, "lrelu", 0.3, out_dtype=E.dec_dtype())
        l2 = self.c2(ctx, l1, out_dtype=E.dec_dtype())
        gamma = self.g(ctx, z, "lrelu", 0.3)
        beta = self.b(ctx, z, "lrelu", 0.3)
        if (l2.data.dtype == l1.data.dtype and gamma.data.dtype == beta.data.dtype == torch.float32
                and l2.data.dtype in (torch.float32, torch.bfloat16)):
            # FiLM -> LeakyReLU -> Add as one pass over the feature map (same arithmetic, same order)
            return E.film_act_add(ctx, l2, gamma, beta, l1, "lrelu", 0.3)
        l2 = self.film(ctx, [l2, gamma, beta])
        l2 = E.activation(ctx, l2, "lrelu", 0.3)
        return E.add(ctx, l1, l2)


def build(conf):
    scope = BuildScope.current()
    a, r = scope.arena, scope.rng
    ca = conf.anatomy_encoder.output_shape[-1]
    H, W = conf.input_shape[0], conf.input_shape[1]
    if conf.decoder_type == "film":
        c0 = E.Conv2D(a, r, "dec_conv0", ca, 8, 3, 1, "same")
        fl = [_FilmLayer(scope, "dec_film%d" % i, conf.num_z) for i in range(1, 5)]
        out = E.Conv2D(a, r, "dec_out", 8, 1, 1, 1, "same", "glorot_normal")
        layers = [c0] + [l for f in fl for l in f.layers()] + [out]

        c0.bf16_grad = True

        def fwd(ctx, anatomy, z):
            l = c0(ctx, anatomy, "lrelu", 0.3, out_dtype=E.dec_dtype())
            for f in fl:
                l = f(ctx, l, z)
            return out(ctx, l, "tanh")
    elif conf.decoder_type == "spade":
        assert H % 32 == 0 and W % 32 == 0, "the SPADE decoder starts at H/32 (decoder.py:68-69)"
        d0 = E.Dense(a, r, "dec_dense", conf.num_z, H * W * 128 // 1024)
        spec = [(128, 128), (128, 128), (128, 128), (128, 64), (64, 32), (32, 16)]
        blocks = [SpadeBlock(scope, "dec_spade%d" % i, ca, fin, fout) for i, (fin, fout) in enumerate(spec)]
        out = E.Conv2D(a, r, "dec_out", 16, 1, 1, 1, "same", "glorot_normal")
        layers = [d0] + [l for b in blocks for l in b.layers()] + [out]

        def fwd(ctx, anatomy, z):
            l = d0(ctx, z)
            l = E.reshape(ctx, l, (l.shape[0], H // 32, W // 32, 128))
            for i, b in enumerate(blocks):
                if i > 0:
                    l = E.upsample2(ctx, l)
                l = b(ctx, anatomy, l)
            return out(ctx, l, "tanh")
    else:
        raise ValueError("Unknown decoder_type value: " + str(conf.decoder_type))
    return Model("Decoder", layers, fwd, [tuple(conf.anatomy_encoder.output_shape), (conf.num_z,)],
                 [(H, W, 1)], scope)

"""Anatomy encoders (reference: model_components/anatomy_encoder.py:13-155).

``build(conf, name)``: one UNet + Conv2D(out_channels, 1, softmax, name='conv_anatomy') + Rounding.
``AnatomyEncoders(modalities).build(conf)``: two private down paths, ONE shared bottleneck /
up path / conv_anatomy (each shared BatchNorm sees the two modalities as two separate batches).
``conf`` is the ``anatomy_encoder`` sub-config (input_shape, out_channels, filters, downsample,
normalise, rounding).
"""
from .. import engine as E
from ..keras_like import BuildScope, Model
from ..models.unet import UNetDown, UNetUp


def _head(scope, f, out_channels):
    return E.Conv2D(scope.arena, scope.rng, "conv_anatomy", f, out_channels, 1, 1, "same")


def _forward(down, up, head, rounding):
    def fwd(ctx, x):
        l, skips = down(ctx, x)
        l = up(ctx, l, skips)
        logits = head(ctx, l)
        return E.softmax(ctx, logits, rounding=bool(rounding))   # softmax + Rounding (STE) fused
    return fwd


def build(conf, name="Enc_Anatomy"):
    scope = BuildScope.current()
    cin = conf.input_shape[-1]
    down = UNetDown(scope, "", cin, conf.filters, conf.downsample, conf.normalise)
    up = UNetUp(scope, "", conf.filters, conf.downsample, conf.normalise)
    head = _head(scope, conf.filters, conf.out_channels)
    out_shape = tuple(conf.input_shape[:-1]) + (conf.out_channels,)
    return Model(name, down.layers() + up.layers() + [head], _forward(down, up, head, conf.rounding),
                 [tuple(conf.input_shape)], [out_shape], scope)


class AnatomyEncoders(object):
    def __init__(self, modalities):
        self.modalities = modalities

    def build(self, conf):
        scope = BuildScope.current()
        cin = conf.input_shape[-1]
        down1 = UNetDown(scope, "enc1_", cin, conf.filters, conf.downsample, conf.normalise)
        down2 = UNetDown(scope, "enc2_", cin, conf.filters, conf.downsample, conf.normalise)
        up = UNetUp(scope, "shared_", conf.filters, conf.downsample, conf.normalise)
        head = _head(scope, conf.filters, conf.out_channels)
        out_shape = tuple(conf.input_shape[:-1]) + (conf.out_channels,)
        shared = up.layers() + [head]
        enc1 = Model("Enc_Anatomy_%s" % self.modalities[0], down1.layers() + shared,
                     _forward(down1, up, head, conf.rounding), [tuple(conf.input_shape)], [out_shape], scope)
        enc2 = Model("Enc_Anatomy_%s" % self.modalities[1], down2.layers() + shared,
                     _forward(down2, up, head, conf.rounding), [tuple(conf.input_shape)], [out_shape], scope)
        return [enc1, enc2]

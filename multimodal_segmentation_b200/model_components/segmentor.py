"""Segmentor (reference: model_components/segmentor.py:9-29): 2x [conv3x3(64, he_normal) -> BN -> ReLU]
-> conv1x1(num_masks+1) softmax."""
from .. import engine as E
from ..keras_like import BuildScope, Model


def build(conf):
    scope = BuildScope.current()
    a, s, r = scope.arena, scope.state, scope.rng
    cin = conf.anatomy_encoder.output_shape[-1]
    c1 = E.Conv2D(a, r, "seg_conv1", cin, 64, 3, 1, "same", "he_normal")
    b1 = E.BatchNorm(a, s, "seg_bn1", 64)
    c2 = E.Conv2D(a, r, "seg_conv2", 64, 64, 3, 1, "same", "he_normal")
    b2 = E.BatchNorm(a, s, "seg_bn2", 64)
    head = E.Conv2D(a, r, "seg_out", 64, conf.num_masks + 1, 1, 1, "same")

    def fwd(ctx, x):
        l = E.conv_bn(ctx, c1, b1, x, "relu", E.feat_dtype())
        l = E.conv_bn(ctx, c2, b2, l, "relu", E.feat_dtype())
        return E.softmax(ctx, head(ctx, l))

    shp = tuple(conf.anatomy_encoder.output_shape)
    return Model("Segmentor", [c1, b1, c2, b2, head], fwd, [shp], [shp[:-1] + (conf.num_masks + 1,)], scope)

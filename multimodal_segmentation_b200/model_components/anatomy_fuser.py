"""Anatomy fuser (reference: model_components/anatomy_fuser.py:12-38): theta = locnet(a1, a2);
a1_deformed = ThinPlateSpline2D(dims, [5,5], channels)([a1, theta]); fused = Maximum([a1_deformed, a2])."""
from .. import engine as E
from ..keras_like import BuildScope, Model
from ..layers import stn_spline
from ..layers.stn_spline import ThinPlateSpline2D


def build(conf):
    scope = BuildScope.current()
    output_shape = tuple(conf.anatomy_encoder.output_shape)
    dims = output_shape[:-1]
    cp = [5, 5]
    channels = conf.anatomy_encoder.out_channels
    locnet = stn_spline.build_locnet(output_shape, output_shape, cp[0] * cp[1] * 2)
    tps = ThinPlateSpline2D(dims, cp, channels)

    def fwd(ctx, anatomy1, anatomy2):
        theta = locnet(ctx, anatomy1, anatomy2)
        deformed = tps(ctx, [anatomy1, theta])
        fused = E.maximum(ctx, deformed, anatomy2)      # tf.maximum: ties -> first input
        return [deformed, fused]

    def fwd_deform(ctx, anatomy1, anatomy2):
        return tps(ctx, [anatomy1, locnet(ctx, anatomy1, anatomy2)])

    m = Model("Anatomy_Fuser", locnet.layers, fwd, [output_shape, output_shape], [output_shape, output_shape], scope)
    m.locnet = locnet
    m.forward_deform = fwd_deform      # first output only (DAFNet trainers discard the fused map)

    def predict_deform_device(a1, a2):
        m._ensure_device()
        return fwd_deform(E.Ctx(None, training=False), E.Var(a1), E.Var(a2)).data
    m.predict_deform_device = predict_deform_device
    return m

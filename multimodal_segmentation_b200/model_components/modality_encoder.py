"""Modality (VAE) encoder (reference: model_components/modality_encoder.py:13-52).

concat(anatomy, image) -> 4x [conv3x3 stride 2 valid (16,32,64,128, he_normal) + LeakyReLU(.3)]
-> Flatten -> Dense(32, he_normal) + LeakyReLU -> z_mean, z_log_var (Dense(num_z)).
The model returns (mu, log_var); the reparameterised sample and the KL output
(utils/sdnet_utils.py:9-21, costs.py:186-189) are produced by ``engine.vae_sample`` so the normal
sample can be injected (parity) -- ``sample(ctx, mu, lv, eps, ...)`` below.
"""
from .. import engine as E
from ..keras_like import BuildScope, Model


def _out_hw(h):
    for _ in range(4):
        h = (h - 3) // 2 + 1
    return h


def build(conf):
    scope = BuildScope.current()
    a, r = scope.arena, scope.rng
    ca = conf.anatomy_encoder.output_shape[-1]
    ci = conf.input_shape[-1]
    convs, c = [], ca + ci
    for i, f in enumerate((16, 32, 64, 128)):
        convs.append(E.Conv2D(a, r, "encm_conv%d" % (i + 1), c, f, 3, 2, "valid", "he_normal"))
        c = f
    flat = _out_hw(conf.input_shape[0]) * _out_hw(conf.input_shape[1]) * 128
    d1 = E.Dense(a, r, "encm_dense", flat, 32, "he_normal")
    z_mean = E.Dense(a, r, "z_mean", 32, conf.num_z)
    z_log_var = E.Dense(a, r, "z_log_var", 32, conf.num_z)

    def trunk(ctx, anatomy, image):
        l = [anatomy, image]          # Concatenate()([anatomy, image]): the first convolution reads the two sources where they lie
        for cv in convs:
            l = cv(ctx, l, "lrelu", 0.3)
        return d1(ctx, l, "lrelu", 0.3)

    def fwd(ctx, anatomy, image):
        l = trunk(ctx, anatomy, image)
        return z_mean(ctx, l), z_log_var(ctx, l)

    def fwd_mu(ctx, anatomy, image):
        return z_mean(ctx, trunk(ctx, anatomy, image))

    m = Model("Enc_Modality", convs + [d1, z_mean, z_log_var], fwd,
              [tuple(conf.anatomy_encoder.output_shape), tuple(conf.input_shape)], [(conf.num_z,), (conf.num_z,)], scope)
    def fwd_lv(ctx, anatomy, image):
        return z_log_var(ctx, trunk(ctx, anatomy, image))

    m.forward_mu = fwd_mu
    m.mu_layers = convs + [d1, z_mean]
    # Enc_Modality_mu = Model(Enc_Modality.inputs, Enc_Modality.get_layer('z_mean').output)   (models/dafnet.py:126)
    m.register_tap("z_mean", fwd_mu, convs + [d1, z_mean], (conf.num_z,))
    m.register_tap("z_log_var", fwd_lv, convs + [d1, z_log_var], (conf.num_z,))
    return m

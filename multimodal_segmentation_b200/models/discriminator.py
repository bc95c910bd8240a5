"""DCGAN-style discriminator with the Spectral kernel regulariser and LS-GAN loss
(reference: models/discriminator.py:9-45).

conv4x4 s2 (f, he_normal) + LeakyReLU(.2); downsample_blocks x [conv4x4 (s2, ..., s1), filters f*2*2^i,
kernel_regularizer=Spectral(f*2^i*16, 10.)] + LeakyReLU(.2); Flatten; Dense(1).
"""
from .. import engine as E
from ..keras_like import BuildScope, Model
from ..layers.spectralnorm import Spectral


class Discriminator(object):
    def __init__(self, conf):
        self.conf = conf
        self.model = None

    def build(self):
        conf = self.conf
        scope = BuildScope.current()
        a, r = scope.arena, scope.rng
        H, W, C = conf.input_shape
        f = conf.filters
        blocks = 3 if not hasattr(conf, "downsample_blocks") else conf.downsample_blocks
        assert blocks > 1, blocks
        convs = [E.Conv2D(a, r, "%s_conv0" % conf.name, C, f, 4, 2, "valid", "he_normal")]
        regs = []
        h, w = (H - 4) // 2 + 1, (W - 4) // 2 + 1
        cin = f
        for i in range(blocks):
            s = 1 if i == blocks - 1 else 2
            cout = f * 2 * (2 ** i)
            cv = E.Conv2D(a, r, "%s_conv%d" % (conf.name, i + 1), cin, cout, 4, s, "valid", "he_normal")
            convs.append(cv)
            regs.append((cv, Spectral(f * (2 ** i) * 4 * 4, 10., r)))
            h, w = (h - 4) // s + 1, (w - 4) // s + 1
            cin = cout
        dense = E.Dense(a, r, "%s_dense" % conf.name, h * w * cin, 1)

        def fwd(ctx, x):
            l = x
            for i, cv in enumerate(convs):      # activated maps between the convolutions in the feature dtype (bf16 on the
                l = cv(ctx, l, "lrelu", 0.2,    # tensor-core path: the next convolution's operand); the last one feeds Dense
                       out_dtype=E.feat_dtype() if i + 1 < len(convs) else None)
            return dense(ctx, l)

        self.model = Model(conf.name, convs + [dense], fwd, [(H, W, C)], [(1,)], scope)
        self.model.regularizers = regs
        return self.model

    def compile(self):
        assert self.model is not None, "Model has not been built"

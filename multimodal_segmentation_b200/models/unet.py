"""UNet building blocks (reference: models/unet.py:16-101, utils/model_utils.py:6-22).

4 down blocks [conv3x3-BN-ReLU]x2 + maxpool, a bottleneck block, 4 up blocks
[Up x2 -> conv3x3 -> BN(linear) -> Concatenate([up, skip]) -> conv block].  On the B200 path the
wide 3x3 convolutions run on tcgen05 with bf16 feature maps written directly by the BN-apply /
pool / upsample kernels; the Concatenate is never materialised (two-source K loop).
"""
from .. import engine as E


def normalise(arena, state, name, c, norm):
    """utils/model_utils.py:6-12 -- 'batch' -> BatchNormalization(); 'instance' -> keras_contrib
    InstanceNormalization() (axis=None, scalar gamma / beta); anything else -> identity."""
    if norm == "batch":
        return E.BatchNorm(arena, state, name, c)
    if norm == "instance":
        return E.InstanceNorm(arena, name)
    return None


class ConvBlock:
    """models/unet.py:94-101 conv_block: [Conv2D(f,3,'same',he_normal) -> norm -> relu] x 2"""

    def __init__(self, scope, name, cin, f, norm):
        a, s, r = scope.arena, scope.state, scope.rng
        self.c1 = E.Conv2D(a, r, name + "_conv1", cin, f, 3, 1, "same", "he_normal")
        self.n1 = normalise(a, s, name + "_bn1", f, norm)
        self.c2 = E.Conv2D(a, r, name + "_conv2", f, f, 3, 1, "same", "he_normal")
        self.n2 = normalise(a, s, name + "_bn2", f, norm)

    def layers(self):
        return [l for l in (self.c1, self.n1, self.c2, self.n2) if l is not None]

    def __call__(self, ctx, x, out_dtype=None):
        fd = E.feat_dtype()
        od = fd if out_dtype is None else out_dtype
        l = E.conv_bn(ctx, self.c1, self.n1, x, "relu", fd) if self.n1 is not None else \
            E.activation(ctx, self.c1(ctx, x), "relu")
        return E.conv_bn(ctx, self.c2, self.n2, l, "relu", od) if self.n2 is not None else \
            E.activation(ctx, self.c2(ctx, l), "relu")


class UpsampleBlock:
    """utils/model_utils.py:15-22 upsample_block(activation='linear'): Up x2 -> conv3x3 -> norm"""

    def __init__(self, scope, name, cin, f, norm):
        a, s, r = scope.arena, scope.state, scope.rng
        self.conv = E.Conv2D(a, r, name + "_conv", cin, f, 3, 1, "same", "he_normal")
        self.norm = normalise(a, s, name + "_bn", f, norm)

    def layers(self):
        return [l for l in (self.conv, self.norm) if l is not None]

    def __call__(self, ctx, x):
        l = E.upsample2(ctx, x)
        if self.norm is None:
            return self.conv(ctx, l)
        return E.conv_bn(ctx, self.conv, self.norm, l, None, E.feat_dtype())


class UNetDown:
    """models/unet.py:37-52 unet_downsample: conv blocks with f, 2f, 4f, 8f filters + 2x2 max pooling"""

    def __init__(self, scope, prefix, cin, f, downsample, norm):
        self.blocks = []
        c = cin
        for i in range(downsample):
            self.blocks.append(ConvBlock(scope, "%sd%d" % (prefix, i), c, f * (2 ** i), norm))
            c = f * (2 ** i)
        self.cout = c

    def layers(self):
        return [l for b in self.blocks for l in b.layers()]

    def __call__(self, ctx, x):
        skips = []
        l = x
        for b in self.blocks:
            d = b(ctx, l)
            skips.append(d)
            l = E.maxpool2(ctx, d)
        return l, skips


class UNetUp:
    """models/unet.py:54-86 bottleneck + unet_upsample (shared between the two modalities in DAFNet,
    model_components/anatomy_encoder.py:100-155)."""

    def __init__(self, scope, prefix, f, downsample, norm):
        fb = f * (2 ** downsample)
        self.bottleneck = ConvBlock(scope, prefix + "bt", f * (2 ** (downsample - 1)), fb, norm)
        self.ups, self.blocks = [], []
        c = fb
        for i in reversed(range(downsample)):
            fi = f * (2 ** i)
            self.ups.append(UpsampleBlock(scope, "%su%d_up" % (prefix, i), c, fi, norm))
            self.blocks.append(ConvBlock(scope, "%su%d" % (prefix, i), 2 * fi, fi, norm))   # Concatenate([up, skip])
            c = fi

    def layers(self):
        out = self.bottleneck.layers()
        for u, b in zip(self.ups, self.blocks):
            out += u.layers() + b.layers()
        return out

    def __call__(self, ctx, l, skips):
        l = self.bottleneck(ctx, l)
        n = len(self.ups)
        for j, (u, b) in enumerate(zip(self.ups, self.blocks)):
            up = u(ctx, l)
            # every block (the last one feeds the 1x1 conv_anatomy on the raster-strip tcgen05 kernel) stores bf16
            l = b(ctx, [up, skips[n - 1 - j]])
        return l

"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/ref_ops.py for the pinning statement).

CPU restatement of the reference's model graphs, written against the reference files (cited per
function) on top of oracle/ref_ops.py.  Weights come in as a dict {name: torch tensor}; names are the
only thing shared with the product (they identify which Keras layer a tensor belongs to).
All randomness (reparameterisation noise, sampled z, spectral u0, fake-sample indices) is injected.
"""
import numpy as np
import torch

from . import ref_ops as R


class BNState(object):
    """collects BatchNorm moving-average updates in call order (a layer shared by k call sites gets
    k updates per step)"""

    def __init__(self, W, training):
        self.W, self.training = W, training
        self.moving = {}

    def get(self, name, key):
        return self.moving.get((name, key), self.W[name + "/" + key].detach())

    def update(self, name, mean, var, count):
        mm, mv = R.bn_moving_update(self.get(name, "moving_mean"), self.get(name, "moving_variance"),
                                    mean.detach(), var.detach(), count)
        self.moving[(name, "moving_mean")] = mm
        self.moving[(name, "moving_variance")] = mv


# ---- optional emulation of the product's tensor-core operand precision ---------------------------------
# The tcgen05 path multiplies bf16 operands (fp32 accumulate) and stores the feature maps that feed it as
# bf16.  A deep random-init ReLU network amplifies ANY perturbation by ~1.3x per layer, so against the
# plain fp32 oracle the 0.3 % bf16 operand noise grows to several % after 23 layers -- for every bf16
# implementation.  With BF16_EMULATION the oracle rounds exactly the tensors the product rounds (operands
# of the eligible 3x3 convolutions, the stored feature maps, the gradients that are kept in bf16), so the
# comparison is again "same arithmetic up to accumulation order".
BF16_EMULATION = False


class _RoundBF16(torch.autograd.Function):
    """forward: round to bf16; backward: the incoming gradient is rounded to bf16 as well (the product
    keeps both the feature map and its gradient in bf16)"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundFwdBF16(torch.autograd.Function):
    """forward: round to bf16; backward: identity"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGradBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def feat(x):
    """a feature map that the product stores as bf16"""
    return _RoundBF16.apply(x) if BF16_EMULATION else x


def raw(x):
    """a convolution output that feeds a BatchNorm: stored as bf16 by the tensor-core convolutions (the batch statistics
    are taken from the rounded values); its gradient comes out of the BatchNorm backward unrounded"""
    return _RoundFwdBF16.apply(x) if BF16_EMULATION else x


def _tc_eligible(w, stride, padding):
    """mirror of the product's eligibility rule (engine.Conv2D.tc_eligible)"""
    if not (BF16_EMULATION and w.shape[0] > 1 and w.shape[2] % 16 == 0 and w.shape[3] % 16 == 0 and w.shape[2] >= 32):
        return False
    return stride == 1 or (stride == 2 and w.shape[0] % 2 == 0 and padding == "valid")


def conv(W, name, x, stride=1, padding="valid"):
    w = W[name + "/kernel"]
    if _tc_eligible(w, stride, padding):
        y = R.conv2d(_RoundBF16.apply(x), _RoundBF16.apply(w), W.get(name + "/bias"), stride, padding)
        return _RoundGradBF16.apply(y)      # the gradient w.r.t. the conv output is consumed as bf16
    if BF16_EMULATION and (stride == 1 or padding == "valid"):
        # narrow layers run on the raster-strip tcgen05 kernels (csrc/conv_nc.cu; stride-2 valid layers through
        # space-to-depth): operands and the incoming output gradient are rounded to bf16 while they are staged
        y = R.conv2d(_RoundFwdBF16.apply(x), _RoundFwdBF16.apply(w), W.get(name + "/bias"), stride, padding)
        return _RoundGradBF16.apply(y)
    return R.conv2d(x, w, W.get(name + "/bias"), stride, padding)


def bn(W, name, x, st):
    """keras BatchNormalization() (utils/model_utils.py:10)"""
    if st.training:
        y, mean, var = R.batchnorm_train(x, W[name + "/gamma"], W[name + "/beta"])
        st.update(name, mean, var, x.shape[0] * x.shape[1] * x.shape[2])
        return y
    return R.batchnorm_infer(x, W[name + "/gamma"], W[name + "/beta"], st.get(name, "moving_mean"),
                             st.get(name, "moving_variance"))


def conv_block(W, name, x, st, last_fp32=False):
    """models/unet.py:94-101"""
    l = feat(R.relu(bn(W, name + "_bn1", raw(conv(W, name + "_conv1", x, 1, "same")), st)))
    l = R.relu(bn(W, name + "_bn2", raw(conv(W, name + "_conv2", l, 1, "same")), st))
    return l if last_fp32 else feat(l)


def upsample_block(W, name, x, st):
    """utils/model_utils.py:15-22 with activation='linear'"""
    return feat(bn(W, name + "_bn", raw(conv(W, name + "_conv", R.upsample2(x), 1, "same")), st))


def anatomy_encoder(W, x, st, down_prefix, up_prefix, downsample=4, rounding=True, head_prefix=""):
    """models/unet.py:37-86 + model_components/anatomy_encoder.py:13-30 (down_prefix == up_prefix == '')
    or :32-98 (private down path, shared up path)."""
    skips = []
    l = x
    for i in range(downsample):
        d = conv_block(W, "%sd%d" % (down_prefix, i), l, st)
        skips.append(d)
        l = R.maxpool2(d)
    l = conv_block(W, up_prefix + "bt", l, st)
    for i in reversed(range(downsample)):
        up = upsample_block(W, "%su%d_up" % (up_prefix, i), l, st)
        l = torch.cat([up, skips[i]], -1)              # Concatenate()([l, self.d_l3]) unet.py:68
        l = conv_block(W, "%su%d" % (up_prefix, i), l, st)      # with BF16_EMULATION every block stores bf16
    a = R.softmax(conv(W, head_prefix + "conv_anatomy", l, 1, "same"))
    return R.rounding(a) if rounding else a


def segmentor(W, s, st):
    """model_components/segmentor.py:9-29"""
    l = feat(R.relu(bn(W, "seg_bn1", raw(conv(W, "seg_conv1", s, 1, "same")), st)))
    l = feat(R.relu(bn(W, "seg_bn2", raw(conv(W, "seg_conv2", l, 1, "same")), st)))
    return R.softmax(conv(W, "seg_out", l, 1, "same"))


def modality_encoder(W, s, x):
    """model_components/modality_encoder.py:34-52 -> (z_mean, z_log_var)"""
    l = torch.cat([s, x], -1)
    for i in range(1, 5):
        l = R.leaky_relu(conv(W, "encm_conv%d" % i, l, 2, "valid"), 0.3)
    l = l.reshape(l.shape[0], -1)                       # Flatten in H,W,C order
    l = R.leaky_relu(R.dense(l, W["encm_dense/kernel"], W["encm_dense/bias"]), 0.3)
    return (R.dense(l, W["z_mean/kernel"], W["z_mean/bias"]),
            R.dense(l, W["z_log_var/kernel"], W["z_log_var/bias"]))


def decoder_film(W, s, z):
    """model_components/decoder.py:36-64 + :28"""
    l = R.leaky_relu(conv(W, "dec_conv0", s, 1, "same"), 0.3)
    for i in range(1, 5):
        n = "dec_film%d" % i
        l1 = R.leaky_relu(conv(W, n + "_conv1", l, 1, "same"), 0.3)
        l2 = conv(W, n + "_conv2", l1, 1, "same")
        gamma = R.leaky_relu(R.dense(z, W[n + "_gamma/kernel"], W[n + "_gamma/bias"]), 0.3)
        beta = R.leaky_relu(R.dense(z, W[n + "_beta/kernel"], W[n + "_beta/bias"]), 0.3)
        l2 = R.leaky_relu(R.film(l2, gamma, beta), 0.3)
        l = l1 + l2
    return torch.tanh(conv(W, "dec_out", l, 1, "same"))


def _spade(W, name, anatomy, layer):
    """layers/spade.py:26-32"""
    layer = R.instance_norm_axis_none(layer)
    an = R.resize_nn(anatomy, layer.shape[1], layer.shape[2])
    an = R.relu(conv(W, name + "_shared", an, 1, "same"))
    return R.spade_cond(layer, conv(W, name + "_gamma", an, 1, "same"), conv(W, name + "_beta", an, 1, "same"))


def spade_block(W, name, anatomy, layer, fin, fout):
    """layers/spade.py:7-23"""
    l = conv(W, name + "_conv1", R.leaky_relu(_spade(W, name + "_s1", anatomy, layer), 0.2), 1, "same")
    l = conv(W, name + "_conv2", R.leaky_relu(_spade(W, name + "_s2", anatomy, l), 0.2), 1, "same")
    sc = layer
    if fin != fout:
        sc = conv(W, name + "_convs", _spade(W, name + "_ss", anatomy, layer), 1, "same")
    return sc + l


def decoder_spade(W, s, z):
    """model_components/decoder.py:67-81 + :28"""
    H, Wd = s.shape[1], s.shape[2]
    l = R.dense(z, W["dec_dense/kernel"], W["dec_dense/bias"]).reshape(-1, H // 32, Wd // 32, 128)
    spec = [(128, 128), (128, 128), (128, 128), (128, 64), (64, 32), (32, 16)]
    for i, (fin, fout) in enumerate(spec):
        if i > 0:
            l = R.upsample2(l)
        l = spade_block(W, "dec_spade%d" % i, s, l, fin, fout)
    return torch.tanh(conv(W, "dec_out", l, 1, "same"))


def decoder(W, s, z, decoder_type="film"):
    return decoder_film(W, s, z) if decoder_type == "film" else decoder_spade(W, s, z)


def locnet(W, a1, a2):
    """layers/stn_spline.py:94-120"""
    l = torch.cat([a1, a2], -1)
    l = R.maxpool2(R.leaky_relu(conv(W, "loc_conv1", l), 0.3))
    l = R.maxpool2(R.leaky_relu(conv(W, "loc_conv2", l), 0.3))
    l = R.leaky_relu(conv(W, "loc_conv3", l), 0.3)
    l = l.reshape(l.shape[0], -1)
    l = torch.tanh(R.dense(l, W["loc_dense1/kernel"], W["loc_dense1/bias"]))
    theta = R.dense(l, W["loc_theta/kernel"], W["loc_theta/bias"])
    return theta.reshape(theta.shape[0], -1, 2)


def anatomy_fuser(W, a1, a2):
    """model_components/anatomy_fuser.py:12-38 -> (a1_deformed, fused, theta)"""
    theta = locnet(W, a1, a2)
    deformed, _ = R.thin_plate_spline_2d(a1, theta, (5, 5), 2, False)
    return deformed, R.tf_maximum(deformed, a2), theta


def discriminator(W, name, x, blocks=3):
    """models/discriminator.py:16-41"""
    l = R.leaky_relu(conv(W, name + "_conv0", x, 2), 0.2)
    for i in range(blocks):
        s = 1 if i == blocks - 1 else 2
        l = R.leaky_relu(conv(W, "%s_conv%d" % (name, i + 1), l, s), 0.2)
    l = l.reshape(l.shape[0], -1)
    return R.dense(l, W[name + "_dense/kernel"], W[name + "_dense/bias"])


def discriminator_trainer_loss(W, name, real, fake, u0s, blocks=3):
    """models/dafnet.py:75-94: mse(D(real),1) + mse(D(fake),0) + the Spectral regularisation losses"""
    lr = R.mse(torch.ones(real.shape[0], 1, dtype=real.dtype), discriminator(W, name, real, blocks))
    lf = R.mse(torch.zeros(fake.shape[0], 1, dtype=fake.dtype), discriminator(W, name, fake, blocks))
    reg = sum(R.spectral_reg(W["%s_conv%d/kernel" % (name, i + 1)], u0s[i], 10.0) for i in range(blocks))
    return lr + lf + reg, (lr, lf, reg)


def dafnet_generator_loss(W, conf, x1, x2, z1_in, z2_in, eps1, eps2, m1, m2=None, supervised=True, training=True):
    """models/dafnet.py:163-222 (graph) + :145-149 (losses, weights) with the targets fed by
    model_executors/dafnet_executor.py:404-410 / :427-433.  Returns (total, dict of per-output losses,
    dict of named intermediate tensors, BNState).  ``training=False`` evaluates the same graph in the inference phase
    (moving BatchNorm statistics): inter["outputs"] is then what the reference trainer's ``predict`` returns
    (tests/test_oracle_builders.py)."""
    st = BNState(W, training=training)
    nm = conf["num_masks"]
    dt = conf.get("decoder_type", "film")
    s1 = anatomy_encoder(W, x1, st, "enc1_", "shared_")
    s2 = anatomy_encoder(W, x2, st, "enc2_", "shared_")
    mu1, lv1 = modality_encoder(W, s1, x1)
    mu2, lv2 = modality_encoder(W, s2, x2)
    z1, kl1 = R.sampling(mu1, lv1, eps1), R.kl(mu1, lv1)
    z2, kl2 = R.sampling(mu2, lv2, eps2), R.kl(mu2, lv2)
    M1 = segmentor(W, s1, st)
    M2 = segmentor(W, s2, st)
    y1 = decoder(W, s1, z1, dt)
    y2 = decoder(W, s2, z2, dt)
    adv_m1 = discriminator(W, "D_Mask", M1[..., 0:nm])
    adv_m2 = discriminator(W, "D_Mask", M2[..., 0:nm])
    adv_y1 = discriminator(W, "D_Image1", y1)
    adv_y2 = discriminator(W, "D_Image2", y2)
    s1_def, _, th1 = anatomy_fuser(W, s1, s2)
    s2_def, _, th2 = anatomy_fuser(W, s2, s1)
    M2_s1_def = segmentor(W, s1_def, st)
    M1_s2_def = segmentor(W, s2_def, st)
    y2_s1_def = decoder(W, s1_def, z2, dt)
    y1_s2_def = decoder(W, s2_def, z1, dt)
    adv_m2_s1_def = discriminator(W, "D_Mask", M2_s1_def[..., 0:nm])
    adv_m1_s2_def = discriminator(W, "D_Mask", M1_s2_def[..., 0:nm])
    adv_y2_s1_def = discriminator(W, "D_Image2", y2_s1_def)
    adv_y1_s2_def = discriminator(W, "D_Image1", y1_s2_def)
    # Z-regressor (dafnet.py:336-350)
    z1_rec = modality_encoder(W, s1, decoder(W, s1, z1_in, dt))[0]
    z2_rec = modality_encoder(W, s2, decoder(W, s2, z2_in, dt))[0]

    ones = lambda t: torch.ones_like(t)
    L = {}
    if supervised:
        seg = [(m1, M1), (m2, M2), (m1, M1_s2_def), (m2, M2_s1_def)]
    else:
        seg = [(m1, M1), (m1, M1_s2_def)]
    for i, (tgt, pred) in enumerate(seg):
        L["Segmentor_%d" % i] = conf["w_sup_M"] * R.combined_dice_bce(tgt, pred, nm)
    for i, a in enumerate((adv_m1, adv_m2, adv_m1_s2_def, adv_m2_s1_def)):
        L["D_Mask_%d" % i] = conf["w_adv_M"] * R.mse(ones(a), a)
    for i, (tgt, y) in enumerate(((x1, y1), (x2, y2), (x1, y1_s2_def), (x2, y2_s1_def))):
        L["Decoder_%d" % i] = conf["w_rec_X"] * R.mae(tgt, y)
    for i, a in enumerate((adv_y1, adv_y2, adv_y1_s2_def, adv_y2_s1_def)):
        L["D_Image_%d" % i] = conf["w_adv_X"] * R.mse(ones(a), a)
    L["KL_0"] = conf["w_kl"] * kl1.mean()
    L["KL_1"] = conf["w_kl"] * kl2.mean()
    L["ZRec_0"] = conf["w_rec_Z"] * R.mae(z1_in, z1_rec)
    L["ZRec_1"] = conf["w_rec_Z"] * R.mae(z2_in, z2_rec)
    total = sum(L.values())
    inter = dict(s1=s1, s2=s2, M1=M1, M2=M2, y1=y1, y2=y2, s1_def=s1_def, s2_def=s2_def, theta1=th1, theta2=th2,
                 z1=z1, z2=z2, mu1=mu1, lv1=lv1, z1_rec=z1_rec, M1_s2_def=M1_s2_def, y1_s2_def=y1_s2_def)
    # the trainer's output list, models/dafnet.py:214-220
    inter["outputs"] = ([M1, M2, M1_s2_def, M2_s1_def] if supervised else [M1, M1_s2_def]) + \
        [adv_m1, adv_m2, adv_m1_s2_def, adv_m2_s1_def] + [y1, y2, y1_s2_def, y2_s1_def] + \
        [adv_y1, adv_y2, adv_y1_s2_def, adv_y2_s1_def] + [kl1, kl2, z1_rec, z2_rec]
    return total, L, inter, st


def balancer(W, s_mod2, s_list):
    """model_components/balancer.py:11-31 + models/dafnet.py:352-361 -> softmax weights [B, n_pairs]"""
    overlap = torch.cat([R.pair_dice(s_mod2, s) for s in s_list], -1)
    l = R.relu(R.dense(overlap, W["bal_dense/kernel"], W["bal_dense/bias"]))
    return R.softmax(R.dense(l, W["beta/kernel"], W["beta/bias"]))


def dafnet_generator_loss_automated(W, conf, x1_lst, x2_lst, z1_in, z2_in, eps1, eps2, m1, m2=None, supervised=True,
                                    training=True):
    """models/dafnet.py:250-334 (get_params_automated_pairing) + :229-235 (losses, weights) with the targets fed by
    model_executors/dafnet_executor.py:447-454 / :470-477.  n_pairs candidate images per modality; candidate 0 is the
    expertly paired one.  Every encoder / segmentor call site is a separate application (own BatchNorm statistics).
    ``training=False``: inference phase; inter["outputs"] is then the reference trainer's ``predict`` output list."""
    st = BNState(W, training=training)
    nm = conf["num_masks"]
    dt = conf.get("decoder_type", "film")
    x1, x2 = x1_lst[0], x2_lst[0]
    s1_lst = [anatomy_encoder(W, x, st, "enc1_", "shared_") for x in x1_lst]
    s2_lst = [anatomy_encoder(W, x, st, "enc2_", "shared_") for x in x2_lst]
    s1, s2 = s1_lst[0], s2_lst[0]
    mu1, lv1 = modality_encoder(W, s1, x1)
    mu2, lv2 = modality_encoder(W, s2, x2)
    z1, kl1 = R.sampling(mu1, lv1, eps1), R.kl(mu1, lv1)
    z2, kl2 = R.sampling(mu2, lv2, eps2), R.kl(mu2, lv2)
    M1 = segmentor(W, s1, st)
    M2 = segmentor(W, s2, st)
    y1 = decoder(W, s1, z1, dt)
    y2 = decoder(W, s2, z2, dt)
    adv_m1 = discriminator(W, "D_Mask", M1[..., 0:nm])
    adv_m2 = discriminator(W, "D_Mask", M2[..., 0:nm])
    adv_y1 = discriminator(W, "D_Image1", y1)
    adv_y2 = discriminator(W, "D_Image2", y2)
    s1_def_lst = [anatomy_fuser(W, s1_i, s2)[0] for s1_i in s1_lst]
    w1 = balancer(W, s2, s1_def_lst)
    s2_def_lst = [anatomy_fuser(W, s2_i, s1)[0] for s2_i in s2_lst]
    w2 = balancer(W, s1, s2_def_lst)
    P = len(x1_lst)
    y2_s1_def_lst = [decoder(W, sd, z2, dt) for sd in s1_def_lst]
    y1_s2_def_lst = [decoder(W, sd, z1, dt) for sd in s2_def_lst]
    y2_s1_def = sum(w1[:, j:j + 1] * R.mae_single_input(x2, y2_s1_def_lst[j]) for j in range(P))
    y1_s2_def = sum(w2[:, j:j + 1] * R.mae_single_input(x1, y1_s2_def_lst[j]) for j in range(P))
    M1_s2_def_lst = [segmentor(W, sd, st) for sd in s2_def_lst]
    # Multiply()([w [B,1], loss [B]]): keras expands the rank-1 loss to [B,1]
    m1_s2_def = sum(w2[:, j:j + 1] * R.combined_dice_bce_perbatch(m1, M1_s2_def_lst[j], nm)[:, None] for j in range(P))
    M2_s1_def_lst = [segmentor(W, sd, st) for sd in s1_def_lst]
    if supervised:
        m2_s1_def = sum(w1[:, j:j + 1] * R.combined_dice_bce_perbatch(m2, M2_s1_def_lst[j], nm)[:, None] for j in range(P))
    adv_m2_s1_def = discriminator(W, "D_Mask", M2_s1_def_lst[0][..., 0:nm])
    adv_m1_s2_def = discriminator(W, "D_Mask", M1_s2_def_lst[0][..., 0:nm])
    adv_y2_s1_def = discriminator(W, "D_Image2", y2_s1_def_lst[0])
    adv_y1_s2_def = discriminator(W, "D_Image1", y1_s2_def_lst[0])
    z1_rec = modality_encoder(W, s1, decoder(W, s1, z1_in, dt))[0]
    z2_rec = modality_encoder(W, s2, decoder(W, s2, z2_in, dt))[0]

    ones = lambda t: torch.ones_like(t)
    L = {}
    if supervised:
        L["Segmentor_0"] = conf["w_sup_M"] * R.combined_dice_bce(m1, M1, nm)
        L["Segmentor_1"] = conf["w_sup_M"] * R.combined_dice_bce(m2, M2, nm)
        L["SegmentorDef_0"] = conf["w_sup_M"] * m1_s2_def.mean()          # costs.ypred
        L["SegmentorDef_1"] = conf["w_sup_M"] * m2_s1_def.mean()
    else:
        L["Segmentor_0"] = conf["w_sup_M"] * R.combined_dice_bce(m1, M1, nm)
        L["SegmentorDef_0"] = conf["w_sup_M"] * m1_s2_def.mean()
    for i, a in enumerate((adv_m1, adv_m2, adv_m1_s2_def, adv_m2_s1_def)):
        L["D_Mask_%d" % i] = conf["w_adv_M"] * R.mse(ones(a), a)
    L["Decoder_0"] = conf["w_rec_X"] * R.mae(x1, y1)
    L["Decoder_1"] = conf["w_rec_X"] * R.mae(x2, y2)
    L["DecoderDef_0"] = conf["w_rec_X"] * y1_s2_def.mean()
    L["DecoderDef_1"] = conf["w_rec_X"] * y2_s1_def.mean()
    for i, a in enumerate((adv_y1, adv_y2, adv_y1_s2_def, adv_y2_s1_def)):
        L["D_Image_%d" % i] = conf["w_adv_X"] * R.mse(ones(a), a)
    L["KL_0"] = conf["w_kl"] * kl1.mean()
    L["KL_1"] = conf["w_kl"] * kl2.mean()
    L["ZRec_0"] = conf["w_rec_Z"] * R.mae(z1_in, z1_rec)
    L["ZRec_1"] = conf["w_rec_Z"] * R.mae(z2_in, z2_rec)
    total = sum(L.values())
    inter = dict(s1=s1, s2=s2, w1=w1, w2=w2, M1=M1, y1=y1, s1_def_lst=s1_def_lst, s2_def_lst=s2_def_lst)
    # the trainer's output list, models/dafnet.py:326-332 (the *_def mask / image entries are the weighted losses)
    inter["outputs"] = ([M1, M2, m1_s2_def, m2_s1_def] if supervised else [M1, m1_s2_def]) + \
        [adv_m1, adv_m2, adv_m1_s2_def, adv_m2_s1_def] + [y1, y2, y1_s2_def, y2_s1_def] + \
        [adv_y1, adv_y2, adv_y1_s2_def, adv_y2_s1_def] + [kl1, kl2, z1_rec, z2_rec]
    return total, L, inter, st


def mmsdnet_generator_loss(W, conf, x1, x2, eps, seg_targets, rec_targets, supervised=True, rounding=True,
                           training=True, return_outputs=False):
    """models/mmsdnet.py:95-192 (unsupervised / supervised trainer graphs and their loss lists) with the targets fed by
    model_executors/mmsdnet_executor.py:257-260 / :287-290.  Two independent UNets (weights enc1_*, enc2_*), the fused
    (Maximum) anatomies are segmented / decoded and trained, dice only, one D_Mask, six KL terms.
    Returns (total, dict of per-output losses); with ``return_outputs`` also the trainer's output list (masks, D_Mask
    scores, reconstructions, KL terms) -- in the inference phase (``training=False``) what the reference trainer's
    ``predict`` returns (tests/test_oracle_builders.py)."""
    st = BNState(W, training=training)
    nm = conf["num_masks"]
    dt = conf.get("decoder_type", "film")
    x = [x1, x2]
    s = [anatomy_encoder(W, x[i], st, "enc%d_" % (i + 1), "enc%d_" % (i + 1), rounding=rounding, head_prefix="enc%d_" % (i + 1))
         for i in range(2)]
    mul = [modality_encoder(W, s[i], x[i]) for i in range(2)]
    z = [R.sampling(mul[i][0], mul[i][1], eps[i]) for i in range(2)]
    kls = [R.kl(*mul[i]) for i in range(2)]
    m1, m2 = segmentor(W, s[0], st), segmentor(W, s[1], st)
    rec = [decoder(W, s[i], z[i], dt) for i in range(2)]
    s1_def, s1_fused, _ = anatomy_fuser(W, s[0], s[1])
    s2_def, s2_fused, _ = anatomy_fuser(W, s[1], s[0])
    fused_seg = [segmentor(W, a, st) for a in (s1_def, s1_fused, s2_def, s2_fused)]
    m_list = ([m1, m2] + fused_seg) if supervised else ([m1] + fused_seg[2:])
    adv_in = [m1, m2] + fused_seg
    for j, a in enumerate((s1_def, s1_fused)):
        mu, lv = modality_encoder(W, a, x[1])
        kls.append(R.kl(mu, lv))
        rec.append(decoder(W, a, R.sampling(mu, lv, eps[2 + j]), dt))
    for j, a in enumerate((s2_def, s2_fused)):
        mu, lv = modality_encoder(W, a, x[0])
        kls.append(R.kl(mu, lv))
        rec.append(decoder(W, a, R.sampling(mu, lv, eps[4 + j]), dt))
    L = {}
    for i, (pred, tgt) in enumerate(zip(m_list, seg_targets)):
        L["Segmentor_%d" % i] = conf["w_sup_M"] * R.dice_loss(tgt, pred, nm)
    adv = []
    for i, m in enumerate(adv_in):
        a = discriminator(W, "D_Mask", m[..., 0:nm])
        adv.append(a)
        L["D_Mask_%d" % i] = conf["w_adv_M"] * R.mse(torch.ones_like(a), a)
    for i, (y, tgt) in enumerate(zip(rec, rec_targets)):
        L["Decoder_%d" % i] = conf["w_rec_X"] * R.mae(tgt, y)
    for i, k in enumerate(kls):
        L["KL_%d" % i] = conf["w_kl"] * k.mean()
    if return_outputs:
        return sum(L.values()), L, m_list + adv + rec + kls
    return sum(L.values()), L


def z_regressor(W, s_list, z_list, decoder_type="film"):
    """models/mmsdnet.py:194-208 / models/dafnet.py:336-350 `build_z_regressor`: every sampled code z_i is decoded on its
    anatomy s_i and re-encoded; returns the list of z_mean(Enc_Modality([s_i, Decoder([s_i, z_i])])) the trainer compares
    with z_i under 'mae' (weight w_rec_Z each)"""
    return [modality_encoder(W, s, decoder(W, s, z, decoder_type))[0] for s, z in zip(s_list, z_list)]


def predict_mask_simple(W, x, down_prefix, up_prefix):
    """models/mmsdnet.py:210-224 type 'simple': Segmentor(Enc_Anatomy(x)) in inference phase"""
    st = BNState(W, training=False)
    return segmentor(W, anatomy_encoder(W, x, st, down_prefix, up_prefix), st)


def predict_mask(W, modality_index, type, x_list):
    """models/mmsdnet.py:210-232, all four types, inference phase: 'simple' Segmentor(s2); 'def' Segmentor(s1 deformed
    onto s2); 'max' Segmentor(Maximum(s1 deformed, s2)); 'maxnostn' Segmentor(np.max([s1, s2]))"""
    assert type in ("simple", "def", "max", "maxnostn")
    st = BNState(W, training=False)
    idx2 = modality_index
    idx1 = 1 - idx2
    pre = ("enc1_", "enc2_")
    s1 = anatomy_encoder(W, x_list[idx1], st, pre[idx1], "shared_")
    s2 = anatomy_encoder(W, x_list[idx2], st, pre[idx2], "shared_")
    if type == "simple":
        return segmentor(W, s2, st)
    if type == "maxnostn":
        return segmentor(W, torch.maximum(s1, s2), st)
    deformed, fused, _ = anatomy_fuser(W, s1, s2)
    return segmentor(W, deformed if type == "def" else fused, st)

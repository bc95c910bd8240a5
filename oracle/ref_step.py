"""ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/ref_ops.py).

One DAFNet `train_batch` on the CPU, following model_executors/dafnet_executor.py:369-583 of the
reference with the oracle graph: generator forward+backward+Adam, two mask-discriminator updates and
two image-discriminator updates, each with its inference-phase fake generation.  Used by bench.py as
the `cpu_baseline` / `--impl reference` arm (TF 1.4 / Keras 2.1.6 cannot be installed here).
"""
import numpy as np
import torch

from . import ref_models as RM
from . import ref_ops as R

_cache = {}


def _weights(conf):
    """Keras-initialised weights, built once on the host by the product's builders (host-only mode is
    not available when CUDA is present, so the numpy initial values are read from the parameter specs)."""
    key = (tuple(conf.input_shape), conf.decoder_type)
    if key in _cache:
        return _cache[key]
    from multimodal_segmentation_b200.keras_like import BuildScope
    from multimodal_segmentation_b200.model_components import anatomy_fuser, decoder, modality_encoder, segmentor
    from multimodal_segmentation_b200.model_components.anatomy_encoder import AnatomyEncoders
    from multimodal_segmentation_b200.models.discriminator import Discriminator
    rng = np.random.RandomState(0)
    W = {}
    with BuildScope(rng=rng) as sc:
        AnatomyEncoders(conf.modality).build(conf.anatomy_encoder)
        anatomy_fuser.build(conf)
        modality_encoder.build(conf)
        segmentor.build(conf)
        decoder.build(conf)
        for name, prm in (("D_Mask", conf.d_mask_params), ("D_Image1", conf.d_image_params), ("D_Image2", conf.d_image_params)):
            prm = dict(prm)
            prm["name"] = name
            from multimodal_segmentation_b200.keras_like import EasyDict
            Discriminator(EasyDict(prm)).build()
    for p in sc.arena.params + sc.state.params:
        W[p.name] = torch.from_numpy(p.init.copy())
    _cache[key] = W
    return W


def _adam(W, names, state, lr=1e-4):
    state["t"] = state.get("t", 0) + 1
    for n in names:
        g = W[n].grad
        if g is None:
            continue
        m, v = state.setdefault(n, (np.zeros(g.shape, np.float32), np.zeros(g.shape, np.float32)))
        p, m, v = R.adam_step(W[n].detach().numpy(), g.numpy(), m, v, state["t"], lr)
        state[n] = (m.astype(np.float32), v.astype(np.float32))
        W[n] = torch.from_numpy(p.astype(np.float32))


def mask_d_candidates(Wd, x1, x2, nm):
    """model_executors/dafnet_executor.py:511-545: the fake masks the two D_Mask updates draw from, BEFORE
    utils.data_utils.sample picks batch_size of them -- modality 1: [Segmentor(s1), Segmentor(Fuser([s2, s1])[0])],
    modality 2: [Segmentor(s2), Segmentor(Fuser([s1, s2])[0])], each cut to the first `nm` channels and concatenated on
    the batch axis; inference phase."""
    inf = RM.BNState(Wd, training=False)
    s1 = RM.anatomy_encoder(Wd, x1, inf, "enc1_", "shared_")
    s2 = RM.anatomy_encoder(Wd, x2, inf, "enc2_", "shared_")
    out = []
    for s_own, s_a, s_b in ((s1, s2, s1), (s2, s1, s2)):
        fm = RM.segmentor(Wd, s_own, inf)
        sdef = RM.anatomy_fuser(Wd, s_a, s_b)[0]
        fmd = RM.segmentor(Wd, sdef, inf)
        out.append(torch.cat([fm[..., :nm], fmd[..., :nm]], 0))
    return out


def image_d_candidates(Wd, x1, x2, eps1, eps2, decoder_type="film"):
    """model_executors/dafnet_executor.py:547-583: the fake images of the two D_Image updates before sampling --
    y1 = [Dec(s1, z1), Dec(s2_def, z1), Dec(s1_def, z1)], y2 = [Dec(s2, z2), Dec(s1_def, z2), Dec(s2_def, z2)] with
    s1_def = Fuser([s1, s2])[0], s2_def = Fuser([s2, s1])[0], z_i the SAMPLED code of Enc_Modality([s_i, x_i])."""
    inf = RM.BNState(Wd, training=False)
    s1 = RM.anatomy_encoder(Wd, x1, inf, "enc1_", "shared_")
    s2 = RM.anatomy_encoder(Wd, x2, inf, "enc2_", "shared_")
    s1d = RM.anatomy_fuser(Wd, s1, s2)[0]
    s2d = RM.anatomy_fuser(Wd, s2, s1)[0]
    mu1, lv1 = RM.modality_encoder(Wd, s1, x1)
    mu2, lv2 = RM.modality_encoder(Wd, s2, x2)
    z1 = R.sampling(mu1, lv1, eps1)
    z2 = R.sampling(mu2, lv2, eps2)
    y1 = torch.cat([RM.decoder(Wd, a, z1, decoder_type) for a in (s1, s2d, s1d)], 0)
    y2 = torch.cat([RM.decoder(Wd, a, z2, decoder_type) for a in (s2, s1d, s2d)], 0)
    return y1, y2


def dafnet_train_batch_cpu(conf, B, seed=0):
    from multimodal_segmentation_b200.loaders.synthetic_chaos import make_pairs
    W = dict(_weights(conf))
    H = conf.input_shape[0]
    nm = conf.num_masks
    rs = np.random.RandomState(seed)
    x1, x2, m1, m2 = make_pairs(B, (H, H, 1), nm, seed=seed)
    res = lambda m: np.concatenate([m, 1 - np.clip(m.sum(-1, keepdims=True), 0, 1)], -1).astype(np.float32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))
    c = dict(num_masks=nm, decoder_type=conf.decoder_type, w_sup_M=conf.w_sup_M, w_adv_M=conf.w_adv_M,
             w_rec_X=conf.w_rec_X, w_adv_X=conf.w_adv_X, w_kl=conf.w_kl, w_rec_Z=conf.w_rec_Z)
    gen_names = [k for k in W if not k.startswith("D_") and "moving_" not in k]
    # ---- generator step (supervised trainer)
    for k in gen_names:
        W[k] = W[k].clone().requires_grad_(True)
    z = [T(rs.normal(size=(B, conf.num_z))) for _ in range(4)]
    total, L, inter, st = RM.dafnet_generator_loss(W, c, T(x1), T(x2), z[0], z[1], z[2], z[3], T(res(m1)), T(res(m2)), True)
    total.backward()
    _adam(W, gen_names, {})
    for (name, key), v in st.moving.items():
        W[name + "/" + key] = v
    with torch.no_grad():
        Wd = {k: v.detach() for k, v in W.items()}
        # ---- mask discriminator x2 (dafnet_executor.py:511-545)
        fakes_m = [cat[rs.choice(2 * B, B, replace=False)] for cat in mask_d_candidates(Wd, T(x1), T(x2), nm)]
        # ---- image discriminators (dafnet_executor.py:547-583)
        e1, e2 = T(rs.normal(size=(B, conf.num_z))), T(rs.normal(size=(B, conf.num_z)))
        c1, c2 = image_d_candidates(Wd, T(x1), T(x2), e1, e2, conf.decoder_type)
        y1 = c1[rs.choice(3 * B, B, replace=False)]
        y2 = c2[rs.choice(3 * B, B, replace=False)]
    for name, real, fake in (("D_Mask", T(m1), fakes_m[0]), ("D_Mask", T(m2), fakes_m[1]),
                             ("D_Image1", T(x1), y1), ("D_Image2", T(x2), y2)):
        names = [k for k in W if k.startswith(name + "_")]
        for k in names:
            W[k] = W[k].detach().clone().requires_grad_(True)
        u0s = [T(rs.uniform(-1, 1, size=(W["%s_conv%d/kernel" % (name, i + 1)].shape[2] * 16, 1))) for i in range(3)]
        loss, _ = RM.discriminator_trainer_loss(W, name, real, fake, u0s)
        loss.backward()
        _adam(W, names, {})
    return float(total)


# ----------------------------------------------------------------------------------------------------------------
# MMSDNet (model_executors/mmsdnet_executor.py:238-331): generator update, Z-regressor update, one mask-discriminator update
# ----------------------------------------------------------------------------------------------------------------
def _mm_enc(Wd, x, i, st, rounding=True):
    p = "enc%d_" % (i + 1)
    return RM.anatomy_encoder(Wd, x, st, p, p, rounding=rounding, head_prefix=p)


def mmsdnet_zreg_anatomies(Wd, x1, x2, rounding=True):
    """mmsdnet_executor.py:266-271: the six anatomies the Z regressor is fitted on, predicted in the inference phase --
    [s1, s2, s1_def, s1_fused, s2_def, s2_fused] with (s1_def, s1_fused) = Fuser([s1, s2]), (s2_def, s2_fused) =
    Fuser([s2, s1])"""
    inf = RM.BNState(Wd, training=False)
    s1, s2 = _mm_enc(Wd, x1, 0, inf, rounding), _mm_enc(Wd, x2, 1, inf, rounding)
    s1_def, s1_fused, _ = RM.anatomy_fuser(Wd, s1, s2)
    s2_def, s2_fused, _ = RM.anatomy_fuser(Wd, s2, s1)
    return [s1, s2, s1_def, s1_fused, s2_def, s2_fused]


def mmsdnet_mask_d_candidates(Wd, x1, x2, nm, rounding=True):
    """mmsdnet_executor.py:319-325: the 4*B fake masks utils.data_utils.sample draws batch_size from --
    [Segmentor(s1), Segmentor(s2), Segmentor(s1_def), Segmentor(s1_fused)] with (s1_def, s1_fused) = Fuser([s1, s2]),
    concatenated on the batch axis and cut to the first `nm` channels; inference phase"""
    inf = RM.BNState(Wd, training=False)
    s1, s2 = _mm_enc(Wd, x1, 0, inf, rounding), _mm_enc(Wd, x2, 1, inf, rounding)
    s1_def, s1_fused, _ = RM.anatomy_fuser(Wd, s1, s2)
    return torch.cat([RM.segmentor(Wd, a, inf)[..., :nm] for a in (s1, s2, s1_def, s1_fused)], 0)


def mmsdnet_zreg_loss(W, conf, s_list, z_list):
    """models/mmsdnet.py:194-208 compiled with 'mae', weight w_rec_Z per output: returns (total, list of the six terms)"""
    zrec = RM.z_regressor(W, s_list, z_list, conf.get("decoder_type", "film"))
    terms = [conf["w_rec_Z"] * R.mae(z, r) for z, r in zip(z_list, zrec)]
    return sum(terms), terms


def _mmsdnet_weights(conf):
    key = ("mmsdnet", tuple(conf.input_shape), conf.decoder_type)
    if key in _cache:
        return _cache[key]
    from multimodal_segmentation_b200.keras_like import BuildScope, EasyDict
    from multimodal_segmentation_b200.model_components import anatomy_encoder, anatomy_fuser, decoder, modality_encoder, segmentor
    from multimodal_segmentation_b200.models.discriminator import Discriminator
    rng = np.random.RandomState(0)
    W = {}
    with BuildScope(rng=rng) as sc:
        encs = [anatomy_encoder.build(conf.anatomy_encoder, "Enc_Anatomy_%s" % m) for m in conf.modality]
        for i, m in enumerate(encs):
            for p in m.weight_list():
                p.name = "enc%d_%s" % (i + 1, p.name)
        anatomy_fuser.build(conf)
        modality_encoder.build(conf)
        segmentor.build(conf)
        decoder.build(conf)
        prm = dict(conf.d_mask_params)
        prm["name"] = "D_Mask"
        Discriminator(EasyDict(prm)).build()
    for p in sc.arena.params + sc.state.params:
        W[p.name] = torch.from_numpy(p.init.copy())
    _cache[key] = W
    return W


def mmsdnet_train_batch_cpu(conf, B, seed=0):
    """one MMSDNet `train_batch` (l_mix = 1) on the CPU: supervised generator update, Z-regressor update, mask-discriminator
    update (mmsdnet_executor.py:238-331).  BASELINE.json config 1 is 66 of these at B = 4."""
    from multimodal_segmentation_b200.loaders.synthetic_chaos import make_pairs
    W = dict(_mmsdnet_weights(conf))
    H = conf.input_shape[0]
    nm = conf.num_masks
    rs = np.random.RandomState(seed)
    x1, x2, m1, m2 = make_pairs(B, (H, H, 1), nm, seed=seed)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))
    c = dict(num_masks=nm, decoder_type=conf.decoder_type, w_sup_M=conf.w_sup_M, w_adv_M=conf.w_adv_M,
             w_rec_X=conf.w_rec_X, w_kl=conf.w_kl, w_rec_Z=conf.w_rec_Z)
    gen_names = [k for k in W if not k.startswith("D_") and "moving_" not in k]
    for k in gen_names:
        W[k] = W[k].clone().requires_grad_(True)
    eps = [T(rs.normal(size=(B, conf.num_z))) for _ in range(6)]
    tm1, tm2, tx1, tx2 = T(m1), T(m2), T(x1), T(x2)
    total, L = RM.mmsdnet_generator_loss(W, c, tx1, tx2, eps, [tm1, tm2, tm2, tm2, tm1, tm1],
                                         [tx1, tx2, tx2, tx2, tx1, tx1], supervised=True)
    total.backward()
    st = {}
    _adam(W, gen_names, st)
    # ---- Z regressor (its own optimizer over Decoder + Enc_Modality; the anatomies are inference-phase predictions)
    with torch.no_grad():
        Wd = {k: v.detach() for k, v in W.items()}
        s_list = mmsdnet_zreg_anatomies(Wd, tx1, tx2)
    z_list = [T(rs.normal(size=(B, conf.num_z))) for _ in range(6)]
    zr_names = [k for k in gen_names if k.startswith("dec_") or k.startswith("encm_") or k.startswith("z_")]
    for k in zr_names:
        W[k] = W[k].detach().clone().requires_grad_(True)
    ztot, _ = mmsdnet_zreg_loss(W, c, s_list, z_list)
    ztot.backward()
    _adam(W, zr_names, {})
    # ---- mask discriminator
    with torch.no_grad():
        Wd = {k: v.detach() for k, v in W.items()}
        fake = mmsdnet_mask_d_candidates(Wd, tx1, tx2, nm)[rs.choice(4 * B, B, replace=False)]
    names = [k for k in W if k.startswith("D_Mask_")]
    for k in names:
        W[k] = W[k].detach().clone().requires_grad_(True)
    u0s = [T(rs.uniform(-1, 1, size=(W["D_Mask_conv%d/kernel" % (i + 1)].shape[2] * 16, 1))) for i in range(3)]
    loss, _ = RM.discriminator_trainer_loss(W, "D_Mask", tm1, fake, u0s)
    loss.backward()
    _adam(W, names, {})
    return float(total.detach())

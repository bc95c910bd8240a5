"""Generate tests/golden/golden_ref.npz by EXECUTING THE REFERENCE'S OWN PYTHON SOURCE (imported unmodified from
/root/reference) on top of the numpy TF/Keras stand-ins of tests/golden/tf_shim.py.

    python tests/golden/make_golden.py          (only runs where /root/reference exists; the .npz is committed)

What is pinned: layers/interpolate_spline.py (solve + apply, orders 1/2/4, regularisation), layers/stn_spline.py
(nDgrid, ThinPlateSpline2D.interpolate_spline_batch / call), layers/film.py, layers/spade.py (SPADE_COND,
resize_like), layers/rounding.py, layers/spectralnorm.py (Spectral.__call__), costs.py (dice, dice losses, the
swapped-argument weighted cross entropy, combined losses, kl, ypred), utils/sdnet_utils.py (sampling),
utils/data_utils.py (rescale, sample), utils/distributions.py, model_executors/base_executor.py (add_residual,
align_batches).  Modules that only exist to reach the data set (loaders, model_tester, image_utils) are stubbed.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import tf_shim  # noqa: E402
from tests.golden.tf_shim import t  # noqa: E402

REF = "/root/reference"


def main():
    out = {}
    rs = np.random.RandomState(1234)
    with tf_shim.installed(REF):
        # modules unrelated to the arithmetic (data set access, plotting) -> empty stubs
        for name in ("utils.image_utils", "loaders", "loaders.loader_factory", "model_tester", "keras_contrib",
                     "keras_contrib.layers"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["loaders"].loader_factory = sys.modules["loaders.loader_factory"]
        sys.modules["model_tester"].ModelTester = object
        sys.modules["keras_contrib.layers"].InstanceNormalization = tf_shim._Dummy
        try:
            from layers import interpolate_spline as ISP
            from layers import stn_spline as STN
            from layers import film as FILM
            from layers import spade as SPADE
            from layers import rounding as RND
            from layers import spectralnorm as SN
            import costs as COSTS
            from utils import sdnet_utils as SDU
            from utils import data_utils as DU
            from utils.distributions import NormalDistribution
            from model_executors import base_executor as BE

            # ---- nDgrid (stn_spline.py:70-91)
            out["ndgrid_5x5"] = np.asarray(STN.nDgrid([5, 5]))
            out["ndgrid_3x4_unnorm"] = np.asarray(STN.nDgrid([3, 4], normalise=False))
            out["ndgrid_3x4_center"] = np.asarray(STN.nDgrid([3, 4], center=True))

            # ---- interpolate_spline (interpolate_spline.py:212-278), the layer's usage: b=1, n=25, d=k=2, order 2
            cp = np.asarray(STN.nDgrid([5, 5]), np.float64)
            q = np.asarray(STN.nDgrid([9, 7]), np.float64)
            theta = rs.normal(size=(3, 25, 2)) * 0.08
            res = [np.asarray(ISP.interpolate_spline(t(cp), t(cp + theta[b:b + 1]), t(q), 2)) for b in range(3)]
            out["spline_cp"], out["spline_q"], out["spline_theta"] = cp, q, theta
            out["spline_order2"] = np.concatenate(res, 0)
            w, v = ISP._solve_interpolation(t(cp), t(cp + theta[0:1]), 2, 0.0)
            out["spline_order2_w0"], out["spline_order2_v0"] = np.asarray(w), np.asarray(v)
            # general scattered points, other orders, regularisation, batch > 1, k = 3
            tp = rs.uniform(size=(2, 11, 2))
            tv = rs.normal(size=(2, 11, 3))
            qq = rs.uniform(size=(2, 17, 2))
            out["spline_tp"], out["spline_tv"], out["spline_qq"] = tp, tv, qq
            out["spline_order1"] = np.asarray(ISP.interpolate_spline(t(tp), t(tv), t(qq), 1))
            out["spline_order4_reg"] = np.asarray(ISP.interpolate_spline(t(tp), t(tv), t(qq), 4, regularization_weight=0.01))
            out["spline_order2_reg"] = np.asarray(ISP.interpolate_spline(t(tp), t(tv), t(qq), 2, regularization_weight=0.003))
            out["spline_order3"] = np.asarray(ISP.interpolate_spline(t(tp), t(tv), t(qq), 3))
            # float32 run of the layer usage (what TF would execute)
            out["spline_order2_f32"] = np.concatenate(
                [np.asarray(ISP.interpolate_spline(t(cp, np.float32), t((cp + theta[b:b + 1]), np.float32), t(q, np.float32), 2))
                 for b in range(3)], 0)

            # ---- ThinPlateSpline2D (stn_spline.py:14-67), inverse False and True
            vol = rs.uniform(size=(3, 9, 7, 4))
            for inv in (False, True):
                layer = STN.ThinPlateSpline2D((9, 7), [5, 5], 4, inverse=inv)
                layer.build(None)
                warped = np.asarray(layer.call([t(vol), t(theta)]))
                locs = np.stack([np.asarray(layer.interpolate_spline_batch(t(theta[b])))[0] for b in range(3)], 0)
                out["tps_warped_inv%d" % inv] = warped
                out["tps_locs_inv%d" % inv] = locs
            out["tps_vol"] = vol

            # ---- FiLM / SPADE_COND / resize_like / Rounding
            x = rs.normal(size=(2, 6, 5, 8))
            g, b_ = rs.normal(size=(2, 8)), rs.normal(size=(2, 8))
            out["film_x"], out["film_gamma"], out["film_beta"] = x, g, b_
            out["film_y"] = np.asarray(FILM.FiLM().call([t(x), t(g), t(b_)]))
            gg, bb = rs.normal(size=x.shape), rs.normal(size=x.shape)
            out["spade_gamma"], out["spade_beta"] = gg, bb
            out["spade_y"] = np.asarray(SPADE.SPADE_COND().call([t(x), t(gg), t(bb)]))
            big = rs.normal(size=(2, 12, 8, 3))
            out["resize_in"] = big
            out["resize_6x4"] = np.asarray(SPADE.resize_like(t(big), t(np.zeros((2, 6, 4, 1)))))
            out["resize_3x2"] = np.asarray(SPADE.resize_like(t(big), t(np.zeros((2, 3, 2, 1)))))
            rx = np.concatenate([rs.uniform(size=40), [0.5, 1.5, 2.5, -0.5, 0.49999997, 0.50000006]]).astype(np.float32)
            out["round_x"] = rx
            out["round_y"] = np.asarray(RND.roundWithGrad(t(rx)))

            # ---- losses (costs.py)
            logits = rs.normal(size=(3, 8, 6, 5))
            pred = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
            lab = rs.randint(0, 5, size=(3, 8, 6))
            true5 = np.eye(5)[lab]
            out["loss_pred"], out["loss_true"] = pred, true5
            out["dice_fnc4"] = np.asarray(COSTS.make_dice_loss_fnc(4)(t(true5), t(pred)))
            out["dice_perbatch4"] = np.asarray(COSTS.dice_coef_perbatch(t(true5[..., :4]), t(pred[..., :4])))
            # keras calls loss(y_true, y_pred); combined_dice_bce forwards them in that order into
            # weighted_cross_entropy_loss(y_pred, y_true) -- the swapped-argument quirk (costs.py:70,134)
            out["wbce_as_called"] = np.asarray(COSTS.weighted_cross_entropy_loss(t(true5), t(pred)))
            out["combined_dice_bce4"] = np.asarray(COSTS.make_combined_dice_bce(4)(t(true5), t(pred)))
            out["combined_perbatch4"] = np.asarray(COSTS.make_combined_dice_bce_perbatch(4)(t(true5[..., :4]), t(logits[..., :4])))
            mu, lv = rs.normal(size=(4, 8)), rs.normal(size=(4, 8)) * 0.3
            out["kl_mu"], out["kl_lv"] = mu, lv
            out["kl"] = np.asarray(COSTS.kl([t(mu), t(lv)]))
            out["ypred"] = np.asarray(COSTS.ypred(None, t(mu)))
            out["dice_metric"] = np.asarray(COSTS.dice(true5[..., :4], pred))
            out["dice_metric_bin"] = np.asarray(COSTS.dice(true5[..., :4], pred, binarise=True))

            # ---- VAE sampling (sdnet_utils.py:9-21); the noise comes from the shim's seeded generator
            tf_shim.RNG["rng"] = np.random.RandomState(7)
            out["sampling_z"] = np.asarray(SDU.sampling([t(mu), t(lv)]))
            out["sampling_eps"] = np.random.RandomState(7).normal(0.0, 1.0, (4, 8))

            # ---- Spectral regulariser (spectralnorm.py:199-246)
            np.random.seed(11)
            reg = SN.Spectral(4 * 4 * 3, 10.)
            out["spectral_u0"] = np.asarray(reg.u).copy()
            Wk = rs.normal(size=(4, 4, 3, 7)) * 0.2
            out["spectral_W"] = Wk
            out["spectral_loss"] = np.asarray(reg(t(Wk))).reshape(())

            # ---- host-side helpers of the step
            img = rs.normal(size=(2, 8, 8, 1))
            out["rescale_in"] = img
            out["rescale_out"] = DU.rescale(img.copy(), -1, 1)
            out["rescale_const"] = DU.rescale(np.full((1, 4, 4, 1), 3.0), -1, 1)
            out["sample_seed5"] = DU.sample(np.arange(20).reshape(10, 2), 4, seed=5)
            m4 = (rs.uniform(size=(2, 6, 6, 4)) > 0.8).astype(np.float64)
            out["residual_in"] = m4
            out["residual_out"] = BE.Executor.add_residual(None, m4)
            al = BE.Executor.align_batches(None, [np.arange(10.).reshape(5, 2), np.arange(6.).reshape(3, 2)])
            out["align_0"], out["align_1"] = al[0], al[1]
            np.random.seed(3)
            out["normal_dist_seed3"] = NormalDistribution().sample((3, 8))

            # ---- automated-pairing losses exactly as models/dafnet.py:283-315 calls them (drawn last: the arrays above
            #      keep their values): SegmentorLoss([m_input, pred]) on 5-channel one-hot masks / softmax outputs,
            #      DecoderLoss([x, y]), and the Balancer's overlap (model_components/balancer.py:33-38)
            out["pb_combined5"] = np.asarray(COSTS.make_combined_dice_bce_perbatch(4)(t(true5), t(pred)))
            out["pb_wce_as_called"] = np.asarray(COSTS.weighted_cross_entropy_perbatch(t(true5), t(pred)))
            xa, ya = rs.uniform(-1, 1, size=(3, 8, 6, 1)), rs.uniform(-1, 1, size=(3, 8, 6, 1))
            out["pb_mae_x"], out["pb_mae_y"] = xa, ya
            out["pb_mae"] = np.asarray(COSTS.mae_single_input([t(xa), t(ya)]))
            from model_components import balancer as BAL
            sa = (rs.uniform(size=(3, 8, 6, 8)) > 0.6).astype(np.float64)
            sb = (rs.uniform(size=(3, 8, 6, 8)) > 0.6).astype(np.float64)
            out["bal_a"], out["bal_b"] = sa, sb
            out["bal_dice"] = np.asarray(BAL.dice([t(sa), t(sb)]))

            # ---- candidate-pair expansion / re-pairing of the automated-pairing data path
            #      (loaders/MultimodalPairedData.py:91-166), the reference class executed from its own file
            #      (skimage's block_reduce, unused here, is the only stand-in); volumes = runs of 8 slices + a short one
            import importlib.util
            for name in ("skimage", "skimage.measure"):
                sys.modules[name] = types.ModuleType(name)
            sys.modules["skimage.measure"].block_reduce = None
            pkg = types.ModuleType("loaders")
            pkg.__path__ = [os.path.join(REF, "loaders")]
            sys.modules["loaders"] = pkg

            def load(name, rel):
                spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
                m = importlib.util.module_from_spec(spec)
                sys.modules[name] = m
                spec.loader.exec_module(m)
                return m

            load("loaders.data", "loaders/data.py")
            MPD = load("loaders.MultimodalPairedData", "loaders/MultimodalPairedData.py").MultimodalPairedData
            n = 19
            imgs = np.concatenate([rs.normal(size=(n, 2, 3, 1)) + 100 * m for m in range(2)], -1).astype(np.float32)
            msks = (rs.uniform(size=(n, 2, 3, 8)) > 0.5).astype(np.float32)
            index = np.arange(n) // 8
            out["pairs_images"], out["pairs_masks"] = imgs, msks
            d = MPD(imgs.copy(), msks.copy(), index)
            np.random.seed(7)
            d.expand_pairs(2, 0, neighborhood=3)
            d.expand_pairs(2, 1, neighborhood=3)
            out["pairs_expand_mod0"], out["pairs_expand_mod1"] = d.get_images_modi(0), d.get_images_modi(1)
            d = MPD(imgs.copy(), msks.copy(), index)
            d.randomise_pairs(length=3, seed=11)
            out["pairs_randomise_images"], out["pairs_randomise_masks"] = d.get_images_modi(0), d.get_masks_modi(0)
            for name in ("skimage", "skimage.measure", "loaders", "loaders.data", "loaders.MultimodalPairedData"):
                sys.modules.pop(name, None)

            # ---- stochastic weight averaging (callbacks/swa.py:27-37): the reference callback driven over 6 epochs
            from callbacks import swa as SWAREF
            hist = [[rs.normal(size=(3, 4)).astype(np.float32), rs.normal(size=(5,)).astype(np.float32)] for _ in range(6)]

            class _Live(object):
                def __init__(self):
                    self.e = 0

                def get_weights(self):
                    return [w.copy() for w in hist[self.e]]

            cb = SWAREF.SWA(2, lambda: None, None)
            cb.model = _Live()
            for e in range(6):
                cb.model.e = e
                cb.on_epoch_end(e)
            for e in range(6):
                out["swa_hist%d_a" % e], out["swa_hist%d_b" % e] = hist[e]
            out["swa_avg_a"], out["swa_avg_b"] = cb.swa_weights
        finally:
            pass
    path = os.path.join(HERE, "golden_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %.1f KB" % (path, len(out), os.path.getsize(path) / 1024.0))
    builders()


def seeded_weights(shapes, seed):
    """weight list drawn from (position, shape) alone: kernels ~ uniform integers scaled to He-like magnitude, biases
    small; float32.  Shared by the generator and tests/test_oracle_builders.py."""
    out = []
    for j, shp in enumerate(shapes):
        shp = tuple(int(v) for v in shp)
        k = np.random.RandomState(seed + j).randint(-32, 33, size=shp).astype(np.float64)
        fan_in = int(np.prod(shp[:-1])) if len(shp) > 1 else 0
        scale = 4.9 / 64.0 / np.sqrt(fan_in) if fan_in else 0.1 / 64.0
        out.append((k * scale).astype(np.float32))
    return out


class _Conf(dict):
    """attribute dictionary (easydict is not installed); the builders only read attributes"""
    __setattr__ = dict.__setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def builders():
    """Run the reference's own BUILDER code (model_components/*.py, models/unet.py, models/discriminator.py,
    layers/stn_spline.build_locnet) on the define-then-run numpy Keras of tests/golden/keras_graph.py and record, per
    component: the inputs, the outputs (inference phase) and the weights in the component's weight order (int8 draws +
    [scale, offset]; value = float32(offset + k * scale)).  tests/test_oracle_builders.py loads the weights into this
    repository's components through set_weights and checks the oracle's restatement of each graph against the outputs."""
    from tests.golden import keras_graph as KG
    out = {}
    rs = np.random.RandomState(4321)
    S = 48                                  # /16 for the UNet, 48 -> 44 -> 22 -> 18 -> 9 -> 5 in the locnet, 48 -> 23 -> 10 -> 4 -> 1 in D
    ae = _Conf(input_shape=(S, S, 1), output_shape=(S, S, 8), out_channels=8, filters=2, downsample=4, normalise="batch",
               rounding=False)
    conf = _Conf(input_shape=(S, S, 1), num_z=8, num_masks=4, decoder_type="film", n_pairs=3, anatomy_encoder=ae)

    B = 1                                   # one sample keeps the committed fixture small

    def anatomy(n=B):
        a = rs.uniform(size=(n, S, S, 8))
        return (a == a.max(-1, keepdims=True)).astype(np.float64)          # one-hot maps, like a rounded softmax

    def f32(a):
        return np.asarray(a, np.float32).astype(np.float64)                # inputs exactly representable in float32

    def record(tag, model, inputs, outputs):
        for i, a in enumerate(inputs):
            a = np.asarray(a)
            out["%s_in%d" % (tag, i)] = a.astype(np.uint8) if np.array_equal(a, a.astype(np.uint8)) else a.astype(np.float32)
        for i, a in enumerate(outputs):
            a = np.asarray(a, np.float32)
            out["%s_out%d" % (tag, i)] = a[:, ::2, ::2] if a.ndim == 4 else a        # maps: every other pixel
        ks = [k for l in model.weighted_layers() for k in l.k]
        sos = [so for l in model.weighted_layers() for so in l.scale_offset]
        out["%s_wk" % tag] = np.concatenate([k.ravel() for k in ks])                  # all integer draws, flat
        out["%s_wshape" % tag] = np.array([list(k.shape) + [0] * (4 - k.ndim) for k in ks], np.int32)
        out["%s_wso" % tag] = np.stack(sos)                                          # [scale, offset] per array

    with tf_shim.installed(REF):
        for name in ("utils.image_utils", "loaders", "loaders.loader_factory", "model_tester", "keras_contrib",
                     "keras_contrib.layers"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["loaders"].loader_factory = sys.modules["loaders.loader_factory"]
        sys.modules["model_tester"].ModelTester = object
        sys.modules["keras_contrib.layers"].InstanceNormalization = tf_shim._Dummy
        sys.modules["callbacks.image_callback"] = types.ModuleType("callbacks.image_callback")     # plotting (matplotlib)
        sys.modules["callbacks.image_callback"].SaveImage = object
        from model_components import segmentor, modality_encoder, decoder, anatomy_fuser, anatomy_encoder, balancer
        from models.discriminator import Discriminator

        KG.reset(101)
        m = segmentor.build(conf)
        x = anatomy()
        record("segmentor", m, [x], [m.predict(x)])

        KG.reset(102)
        m = modality_encoder.build(conf)
        xs = [anatomy(), f32(rs.uniform(-1, 1, size=(B, S, S, 1)))]
        head = KG.Model(inputs=m.inputs, outputs=[m.get_layer("z_mean").output, m.get_layer("z_log_var").output,
                                                  m.get_layer("divergence").output])
        record("modality_encoder", m, xs, head.predict(xs))          # the sampled z itself is random

        KG.reset(103)
        m = decoder.build(conf)
        xs = [anatomy(), f32(rs.normal(size=(B, 8)))]
        record("decoder_film", m, xs, [m.predict(xs)])

        KG.reset(104)
        m = anatomy_fuser.build(conf)
        xs = [anatomy(), anatomy()]
        theta = KG.Model(inputs=m.inputs, outputs=m.get_layer("stn_locnet").output).predict(xs)
        record("anatomy_fuser", m, xs, list(m.predict(xs)) + [theta])

        KG.reset(105)
        m = anatomy_encoder.build(ae)
        x = f32(rs.uniform(-1, 1, size=(B, S, S, 1)))
        record("anatomy_encoder", m, [x], [m.predict(x)])

        KG.reset(106)
        e1, e2 = anatomy_encoder.AnatomyEncoders(["t1", "t2"]).build(ae)
        x1, x2 = f32(rs.uniform(-1, 1, size=(B, S, S, 1))), f32(rs.uniform(-1, 1, size=(B, S, S, 1)))
        y1, y2 = e1.predict(x1), e2.predict(x2)
        record("anatomy_encoders_1", e1, [x1], [y1])
        record("anatomy_encoders_2", e2, [x2], [y2])

        KG.reset(107)
        m = Discriminator(_Conf(input_shape=(S, S, 4), name="D_Mask", filters=4, lr=1e-4)).build()
        x = f32(rs.uniform(size=(B, S, S, 4)))
        record("discriminator", m, [x], [m.predict(x)])

        KG.reset(108)
        m = balancer.build(conf)
        xs = [anatomy(2) for _ in range(4)]
        record("balancer", m, xs, [m.predict(xs)])
        # ---- the generator-step graph itself (models/dafnet.py:140-222,336-350): the reference DAFNet class builds its
        #      components and wires the supervised trainer; its 20 outputs in the inference phase, with the reparametrisation
        #      noise fixed to one array
        from models.dafnet import DAFNet
        import keras.backend as K
        ae_r = _Conf(ae, rounding=True)
        dconf = _Conf(conf, n_pairs=1, automatedpairing=False, modality=["t1", "t2"], anatomy_encoder=ae_r,
                      d_mask_params=_Conf(filters=4, lr=1e-4, name="D_Mask", input_shape=(S, S, 4)),
                      d_image_params=_Conf(filters=4, lr=1e-4, name="D_Image", input_shape=(S, S, 1)),
                      w_sup_M=10, w_adv_M=1, w_rec_X=1, w_adv_X=1, w_rec_Z=1, w_kl=0.1, lr=1e-4, folder="/tmp/dafk_no_such_folder")
        KG.reset(201)
        np.random.seed(2010)                 # layers/spectralnorm.py:213 draws its power-iteration vectors from numpy's RNG
        net = DAFNet(dconf)
        net.loader = _Conf(num_masks=4)
        net.build()
        xs = [f32(rs.uniform(-1, 1, size=(B, S, S, 1))), f32(rs.uniform(-1, 1, size=(B, S, S, 1))),
              f32(rs.normal(size=(B, 8))), f32(rs.normal(size=(B, 8)))]
        eps = f32(rs.normal(size=(B, 8)))
        K.random_normal = lambda shape, mean=0.0, stddev=1.0: t(eps.copy())
        outs = net.supervised_trainer.predict(xs)
        assert len(outs) == 20
        for i, a in enumerate(xs + [eps]):
            out["trainer_in%d" % i] = a.astype(np.float32)
        for i, a in enumerate(outs):
            a = np.asarray(a, np.float32)
            out["trainer_out%02d" % i] = a[:, ::2, ::2] if a.ndim == 4 else a       # big maps: every other pixel
        # the compiled losses and weights (models/dafnet.py:145-149) on the targets the executor feeds
        # (model_executors/dafnet_executor.py:404-410): masks, ones for the discriminator scores, the images, a dummy for
        # the KL outputs (costs.ypred ignores it), the sampled z for the regressor
        lab1, lab2 = np.eye(5)[rs.randint(0, 5, size=(B, S, S))], np.eye(5)[rs.randint(0, 5, size=(B, S, S))]
        out["trainer_m1"], out["trainer_m2"] = lab1.astype(np.uint8), lab2.astype(np.uint8)
        ones, zero = np.ones((B, 1)), np.zeros((B, 1))
        tg = [lab1, lab2, lab1, lab2] + [ones] * 4 + [xs[0], xs[1], xs[0], xs[1]] + [ones] * 4 + [zero, zero, xs[2], xs[3]]
        out["trainer_loss"] = np.array(net.supervised_trainer.loss_values(xs, tg))
        tgu = [lab1, lab1] + tg[4:]
        out["trainer_unsup_loss"] = np.array(net.unsupervised_trainer.loss_values(xs, tgu))
        # the mask-discriminator trainer (models/mmsdnet.py:62-78): mse on D(real) / D(fake) against ones / zeros
        # (dafnet_executor.py:530-534) plus, as Keras adds them, the Spectral regularisers of its convolutions evaluated on
        # their kernels from the initial power-iteration vectors
        real_m = lab1[..., :4]
        fake_m = np.repeat(xs[0], 4, -1) * 0.5 + 0.5
        dvals = net.D_Mask_trainer.loss_values([real_m, fake_m], [ones, zero])
        regs = []
        for j, l in enumerate(net.D_Mask.regularized_layers()):
            out["dtrain_u0_%d" % j] = np.asarray(l.kernel_regularizer.u, np.float64).copy()
            regs.append(float(np.asarray(l.kernel_regularizer(t(l.kernel)))))
        out["dtrain_loss"] = np.array(dvals + regs)
        # ---- the executor's discriminator steps (model_executors/dafnet_executor.py:511-583), run unmodified on this
        #      network: which predictions are concatenated, in which order, before utils.data_utils.sample draws the fakes.
        #      `fit` of the stand-in records what it is fed.
        import itertools
        for name, attr in (("callbacks.dafnet_image_callback", "DAFNetImageCallback"), ("callbacks.loss_callback", "SaveLoss")):
            sys.modules[name] = types.ModuleType(name)
            setattr(sys.modules[name], attr, object)
        from model_executors.dafnet_executor import DAFNetExecutor
        ex = object.__new__(DAFNetExecutor)
        ex.conf, ex.model, ex.loader = dconf, net, _Conf(num_masks=4)
        dx = [f32(rs.uniform(-1, 1, size=(2, S, S, 1))) for _ in range(2)]
        dm = [np.eye(5)[rs.randint(0, 5, size=(2, S, S))] for _ in range(2)]
        out["dstep_x1"], out["dstep_x2"] = dx[0].astype(np.float32), dx[1].astype(np.float32)
        out["dstep_m1"], out["dstep_m2"] = dm[0].astype(np.uint8), dm[1].astype(np.uint8)
        ex.discriminator_masks = itertools.cycle(dm)
        ex.discriminator_image = [itertools.cycle([dx[0]]), itertools.cycle([dx[1]])]
        from collections import defaultdict
        np.random.seed(31)
        ex.train_batch_mask_discriminator(defaultdict(list))
        np.random.seed(32)
        ex.train_batch_image_discriminator(defaultdict(list))
        (r1, f1), _ = net.D_Mask_trainer.fit_calls[-2]
        (r2, f2), _ = net.D_Mask_trainer.fit_calls[-1]
        out["dstep_mask_real1"], out["dstep_mask_real2"] = r1.astype(np.uint8), r2.astype(np.uint8)
        out["dstep_mask_fake1"], out["dstep_mask_fake2"] = f1.astype(np.float32), f2.astype(np.float32)
        (xr1, y1f), tg1 = net.D_Image1_trainer.fit_calls[-1]
        (xr2, y2f), tg2 = net.D_Image2_trainer.fit_calls[-1]
        assert np.array_equal(xr1, dx[0]) and np.array_equal(xr2, dx[1])
        assert np.all(tg1[0] == 1) and np.all(tg1[1] == 0)
        out["dstep_img_fake1"], out["dstep_img_fake2"] = y1f.astype(np.float32), y2f.astype(np.float32)

        # ---- the generator steps of the executor, run unmodified (dafnet_executor.py:389-432,482-500): the targets it
        #      feeds must be the convention used for `tg` / `tgu` above (the oracle's loss functions take the same)
        gx1, gx2 = f32(rs.uniform(-1, 1, size=(2, S, S, 1))), f32(rs.uniform(-1, 1, size=(2, S, S, 1)))
        gm1 = np.eye(5)[rs.randint(0, 5, size=(2, S, S))][..., :4]
        gm2 = np.eye(5)[rs.randint(0, 5, size=(2, S, S))][..., :4]
        ex.gen_labelled = itertools.cycle([(gx1, gx2, gm1, gm2)])
        ex.gen_unlabelled = itertools.cycle([(gx1, gx2, gm1)])
        np.random.seed(33)
        ex.train_supervised_expert_pairing(defaultdict(list))
        (i_x1, i_x2, i_z1, i_z2), tgs = net.supervised_trainer.fit_calls[-1]
        res = lambda m: np.concatenate([m, 1 - m.sum(-1, keepdims=True)], -1)       # base_executor.add_residual on one-hot
        np.random.seed(33)
        z1_expect, z2_expect = np.random.normal(0, 1, (2, 8)), np.random.normal(0, 1, (2, 8))
        assert np.array_equal(i_x1, gx1) and np.array_equal(i_x2, gx2)
        assert np.array_equal(i_z1, z1_expect) and np.array_equal(i_z2, z2_expect)
        expect = [res(gm1), res(gm2), res(gm1), res(gm2)] + [np.ones((2, 1))] * 4 + [gx1, gx2, gx1, gx2] + \
            [np.ones((2, 1))] * 4 + [np.zeros(2), np.zeros(2), z1_expect, z2_expect]
        assert len(tgs) == 20 and all(np.array_equal(a, b) for a, b in zip(tgs, expect))
        ex.train_unsupervised_expert_pairing(defaultdict(list))
        _, tgs_u = net.unsupervised_trainer.fit_calls[-1]
        assert len(tgs_u) == 18 and np.array_equal(tgs_u[0], res(gm1)) and np.array_equal(tgs_u[1], res(gm1))
        assert all(np.array_equal(a, b) for a, b in zip(tgs_u[2:14], expect[4:16]))
        out["executor_targets_checked"] = np.array(1)
        # the step schedule itself (dafnet_executor.py:369-387): which trainer is fitted in which order by train_batch
        net.supervised_trainer.name, net.unsupervised_trainer.name = "supervised_trainer", "unsupervised_trainer"
        for lm in (1, 0.5, 0):
            KG.STATE["fit_log"] = []
            ex.conf = _Conf(dconf, l_mix=lm)
            ex.train_batch(defaultdict(list))
            out["schedule_l_mix_%s" % lm] = np.array(KG.STATE["fit_log"])
        ex.conf = dconf
        # inference entry point (models/mmsdnet.py:210-232, inherited by DAFNet): all four fusion types
        for mi, types_ in ((1, ("simple", "def", "max", "maxnostn")), (0, ("simple", "def"))):
            for ty in types_:
                pm = net.predict_mask(mi, ty, [xs[0], xs[1]])
                out["predict_mask_%d_%s" % (mi, ty)] = np.asarray(pm, np.float32)[:, ::2, ::2]
        # the unsupervised trainer is the same graph with two mask outputs less: store which supervised output each of its
        # 18 outputs equals
        uouts = net.unsupervised_trainer.predict(xs)
        assert len(uouts) == 18
        idx = []
        for u in uouts:
            hit = [j for j, a in enumerate(outs) if np.shape(a) == np.shape(u) and np.array_equal(a, u)]
            assert len(hit) >= 1
            idx.append(hit[0])
        out["trainer_unsup_index"] = np.array(idx)
        for tag, m in (("enc1", net.Encoders_Anatomy[0]), ("enc2", net.Encoders_Anatomy[1]), ("encm", net.Enc_Modality),
                       ("fuser", net.Anatomy_Fuser), ("seg", net.Segmentor), ("dec", net.Decoder), ("dmask", net.D_Mask),
                       ("dimg1", net.D_Image1), ("dimg2", net.D_Image2)):
            record("trainer_" + tag, m, [], [])
        # ---- the automated-pairing trainer (models/dafnet.py:224-334,352-361): three candidate images per modality,
        #      Balancer-weighted per-sample losses computed INSIDE the graph.  Same construction sequence and seed as
        #      above, hence the same component weights (checked); only the Balancer's are new.
        first = {tag: [w.copy() for w in m.get_weights()] for tag, m in (("enc1", net.Encoders_Anatomy[0]), ("seg", net.Segmentor),
                                                                        ("dimg2", net.D_Image2))}
        aconf = _Conf(dconf, n_pairs=3, automatedpairing=True, input_shape=[S, S, 1])
        KG.reset(201)
        anet = DAFNet(aconf)
        anet.loader = _Conf(num_masks=4)
        anet.build()

        def masks():
            lab = rs.randint(0, 5, size=(B, S, S))
            return np.eye(5)[lab]

        axs = [f32(rs.uniform(-1, 1, size=(B, S, S, 1))) for _ in range(6)] + [masks(), masks()] + \
            [f32(rs.normal(size=(B, 8))), f32(rs.normal(size=(B, 8)))]
        aouts = anet.supervised_trainer.predict(axs)
        assert len(aouts) == 20
        for tag, m in (("enc1", anet.Encoders_Anatomy[0]), ("seg", anet.Segmentor), ("dimg2", anet.D_Image2)):
            assert all(np.array_equal(a, b) for a, b in zip(first[tag], m.get_weights())), tag
        for i, a in enumerate(axs):
            out["auto_in%d" % i] = a.astype(np.uint8) if np.array_equal(a, a.astype(np.uint8)) else a.astype(np.float32)
        for i, a in enumerate(aouts):
            a = np.asarray(a, np.float32)
            out["auto_out%02d" % i] = a[:, ::2, ::2] if a.ndim == 4 else a
        record("auto_balancer", anet.Balancer, [], [])
        # targets of the automated-pairing step (dafnet_executor.py:447-454): the *_Def outputs are losses already (ypred)
        atg = [axs[6], axs[7], zero, zero] + [ones] * 4 + [axs[0], axs[3], zero, zero] + [ones] * 4 + [zero, zero, axs[8], axs[9]]
        out["auto_loss"] = np.array(anet.supervised_trainer.loss_values(axs, atg))
        # unsupervised variant (no m2 input, 18 outputs; dafnet_executor.py:470-477)
        uaxs = axs[:7] + axs[8:]
        uatg = [axs[6], zero] + atg[4:]
        out["auto_unsup_loss"] = np.array(anet.unsupervised_trainer.loss_values(uaxs, uatg))
        # ---- SPADE decoder (model_components/decoder.py:67-81, layers/spade.py:7-55).  4.3 M parameters: instead of
        #      storing them, every array of the model's weight list is re-drawn from its position and shape
        #      (seeded_weights below); the test draws the same list for this repository's component
        sys.modules["keras_contrib.layers"].InstanceNormalization = KG.InstanceNormalization
        import importlib
        from layers import spade as SPADE_MOD
        importlib.reload(SPADE_MOD)                      # picks up the InstanceNormalization stand-in
        importlib.reload(decoder)
        sconf = _Conf(conf, decoder_type="spade", input_shape=(64, 64, 1),
                      anatomy_encoder=_Conf(ae, input_shape=(64, 64, 1), output_shape=(64, 64, 8)))
        KG.reset(401)
        m = decoder.build(sconf)
        a64 = rs.uniform(size=(1, 64, 64, 8))
        sx = [(a64 == a64.max(-1, keepdims=True)).astype(np.float64), f32(rs.normal(size=(1, 8)))]
        m.predict(sx)                                    # creates the layers' weights (shapes)
        m.set_weights(seeded_weights([w.shape for w in m.get_weights()], 4010))
        out["decoder_spade_in0"], out["decoder_spade_in1"] = sx[0].astype(np.uint8), sx[1].astype(np.float32)
        out["decoder_spade_out0"] = np.asarray(m.predict(sx), np.float32)[:, ::2, ::2]
        out["decoder_spade_nw"] = np.array(len(m.get_weights()))

        # ---- MMSDNet (models/mmsdnet.py:62-192): two independent anatomy encoders, one D_Mask, the deformed AND the fused
        #      anatomies segmented / re-encoded / decoded; supervised trainer, 24 outputs
        from models.mmsdnet import MMSDNet
        KG.reset(301)
        mnet = MMSDNet(dconf)
        mnet.loader = _Conf(num_masks=4)
        mnet.build()
        mouts = mnet.supervised_trainer.predict([xs[0], xs[1]])
        assert len(mouts) == 24
        for i, a in enumerate(mouts):
            a = np.asarray(a, np.float32)
            out["mmsd_out%02d" % i] = a[:, ::2, ::2] if a.ndim == 4 else a
        for tag, m in (("enc1", mnet.Encoders_Anatomy[0]), ("enc2", mnet.Encoders_Anatomy[1]), ("encm", mnet.Enc_Modality),
                       ("fuser", mnet.Anatomy_Fuser), ("seg", mnet.Segmentor), ("dec", mnet.Decoder), ("dmask", mnet.D_Mask)):
            record("mmsd_" + tag, m, [], [])
        # the six-input Z regressor (models/mmsdnet.py:194-208): anatomies = one stored map with its channels rolled
        za = anatomy(1)
        out["mmsd_zreg_s"] = za.astype(np.uint8)
        zs = [f32(rs.normal(size=(1, 8))) for _ in range(6)]
        out["mmsd_zreg_z"] = np.concatenate(zs, 0).astype(np.float32)
        zr = mnet.Z_Regressor.predict([np.roll(za, i, axis=-1) for i in range(6)] + zs)
        out["mmsd_zreg_out"] = np.concatenate(zr, 0).astype(np.float32)
        out["mmsd_zreg_loss"] = np.array(mnet.Z_Regressor.loss_values([np.roll(za, i, axis=-1) for i in range(6)] + zs, zs))
        # loss list of the supervised trainer (models/mmsdnet.py:181-190) on the executor's targets
        # (model_executors/mmsdnet_executor.py:257-260): masks of modality 1, 2, then 2, 2, 1, 1 for the deformed / fused
        # anatomies; images x1, x2, x2, x2, x1, x1
        mtg = [lab1[..., :4], lab2[..., :4], lab2[..., :4], lab2[..., :4], lab1[..., :4], lab1[..., :4]] + [ones] * 6 + \
            [xs[0], xs[1], xs[1], xs[1], xs[0], xs[0]] + [zero] * 6
        out["mmsd_loss"] = np.array(mnet.supervised_trainer.loss_values([xs[0], xs[1]], mtg))
        # ---- the MMSDNet executor's step (model_executors/mmsdnet_executor.py:238-331), run UNMODIFIED on this network with
        #      recording trainers: which anatomies the Z regressor is fitted on, which fake masks the single D_Mask update
        #      draws from, and the trainer order of train_batch for l_mix in {1, 0.5, 0}
        from model_executors.mmsdnet_executor import MMSDNetExecutor
        mex = object.__new__(MMSDNetExecutor)
        mex.conf, mex.model, mex.loader = dconf, mnet, _Conf(num_masks=4)
        mx = [f32(rs.uniform(-1, 1, size=(2, S, S, 1))) for _ in range(2)]
        mm = [np.eye(5)[rs.randint(0, 5, size=(2, S, S))] for _ in range(2)]
        out["mstep_x1"], out["mstep_x2"] = mx[0].astype(np.float32), mx[1].astype(np.float32)
        out["mstep_m1"], out["mstep_m2"] = mm[0].astype(np.uint8), mm[1].astype(np.uint8)
        mex.discriminator_masks = itertools.cycle([mm[0]])
        mex.discriminator_image = [itertools.cycle([mx[0]]), itertools.cycle([mx[1]])]
        np.random.seed(41)
        mex.train_batch_mask_discriminator(defaultdict(list))
        (mr, mf), mtg_d = mnet.D_Mask_trainer.fit_calls[-1]
        assert np.array_equal(mr, mm[0][..., :4]) and np.all(mtg_d[0] == 1) and np.all(mtg_d[1] == 0)
        out["mstep_mask_fake"] = mf.astype(np.float32)
        mex.gen_labelled = itertools.cycle([(mx[0], mx[1], mm[0][..., :4], mm[1][..., :4])])
        mex.gen_unlabelled = itertools.cycle([(mx[0], mx[1], mm[0][..., :4])])
        mnet.supervised_trainer.name, mnet.unsupervised_trainer.name = "supervised_trainer", "unsupervised_trainer"
        mnet.Z_Regressor.name = "Z_Regressor"
        for lm in (1, 0.5, 0):
            KG.STATE["fit_log"] = []
            mex.conf = _Conf(dconf, l_mix=lm)
            np.random.seed(42)
            mex.train_batch(defaultdict(list))
            out["mmsd_schedule_l_mix_%s" % lm] = np.array(KG.STATE["fit_log"])
        mex.conf = _Conf(dconf, l_mix=1)
        np.random.seed(43)
        mex.train_batch_generators(defaultdict(list))
        (g_in, g_tg) = mnet.supervised_trainer.fit_calls[-1]
        m14a, m14b = mm[0][..., :4], mm[1][..., :4]
        exp_tg = [m14a, m14b, m14b, m14b, m14a, m14a] + [np.ones((2, 1))] * 6 + [mx[0], mx[1], mx[1], mx[1], mx[0], mx[0]] + \
            [np.zeros(2)] * 6
        assert len(g_in) == 2 and np.array_equal(g_in[0], mx[0]) and np.array_equal(g_in[1], mx[1])
        assert len(g_tg) == 24 and all(np.array_equal(a, b) for a, b in zip(g_tg, exp_tg))
        (z_in, z_tg) = mnet.Z_Regressor.fit_calls[-1]
        assert len(z_in) == 12 and len(z_tg) == 6 and all(np.array_equal(a, b) for a, b in zip(z_in[6:], z_tg))
        for i in range(6):
            out["mstep_zreg_s%d" % i] = np.asarray(z_in[i], np.float32)[:, ::2, ::2]
        out["mmsd_executor_targets_checked"] = np.array(1)
    path = os.path.join(HERE, "golden_builders.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %.1f KB" % (path, len(out), os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()

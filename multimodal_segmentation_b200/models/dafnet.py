"""DAFNet (reference: models/dafnet.py:18-361): anatomy encoders with a shared decoder path, modality
encoder, anatomy fuser (locnet + thin-plate-spline STN), segmentor, FiLM/SPADE decoder, three LS-GAN
discriminators; expert-pairing trainers (dafnet.py:140-222), automated-pairing trainers with the Balancer
(dafnet.py:224-334,352-361) and the Z-regressor (dafnet.py:336-350)."""
import logging

from .. import engine as E
from ..keras_like import BuildScope
from ..model_components import anatomy_fuser, balancer, decoder, modality_encoder, segmentor
from ..model_components.anatomy_encoder import AnatomyEncoders
from .discriminator import Discriminator
from .mmsdnet import MMSDNet, MuModel, to_eps
from .trainers import DiscriminatorTrainer, Trainer

log = logging.getLogger("dafnet")


class DAFNetGeneratorTrainer(Trainer):
    """get_params_expert_pairing (dafnet.py:163-222) + the loss dict / weights of
    build_trainers_expertpairs (dafnet.py:145-149).

    graph inputs : x1, x2, z1_in, z2_in, eps1, eps2, m1[, m2]
    outputs/loss : Segmentor x4 (x2 unsupervised): dice + .01*wBCE, weight w_sup_M
                   D_Mask x4: mse vs 1, w_adv_M       Decoder x4: mae, w_rec_X
                   D_Image1/2 x4: mse vs 1, w_adv_X    Enc_Modality x2: mean(KL), w_kl
                   ZReconstruct x2: mae, w_rec_Z
    """

    def __init__(self, net, supervised):
        c = net.conf
        self.net, self.supervised = net, supervised
        nseg = 4 if supervised else 2
        self.nseg = nseg
        nw = [("Segmentor", c.w_sup_M)] * nseg + [("D_Mask", c.w_adv_M)] * 4 + [("Decoder", c.w_rec_X)] * 4 + \
             [("D_Image1", c.w_adv_X), ("D_Image2", c.w_adv_X), ("D_Image1", c.w_adv_X), ("D_Image2", c.w_adv_X)] + \
             [("Enc_Modality", c.w_kl)] * 2 + [("ZReconstruct", c.w_rec_Z)] * 2
        super().__init__("supervised_trainer" if supervised else "unsupervised_trainer", net.generator_params(), c.lr, nw,
                         frozen_models=[net.D_Mask, net.D_Image1, net.D_Image2])

    def graph(self, ctx, book, x1, x2, z1_in, z2_in, eps1, eps2, m1, m2=None):
        n, c = self.net, self.net.conf
        ns = self.nseg
        nm = n.loader.num_masks
        X1, X2 = E.Var(x1), E.Var(x2)
        # encode
        s1 = n.Encoders_Anatomy[0](ctx, X1)
        s2 = n.Encoders_Anatomy[1](ctx, X2)
        mu1, lv1 = n.Enc_Modality(ctx, s1, X1)
        mu2, lv2 = n.Enc_Modality(ctx, s2, X2)
        z1, _ = E.vae_sample(ctx, mu1, lv1, eps1, c.w_kl, book.slot(ns + 12))
        z2, _ = E.vae_sample(ctx, mu2, lv2, eps2, c.w_kl, book.slot(ns + 13))
        # segment / decode
        M1 = n.Segmentor(ctx, s1)
        M2 = n.Segmentor(ctx, s2)
        # deform (the fused output is discarded by the DAFNet trainers, dafnet.py:193-194)
        s1_def = n.Anatomy_Fuser.forward_deform(ctx, s1, s2)
        s2_def = n.Anatomy_Fuser.forward_deform(ctx, s2, s1)
        M2_s1_def = n.Segmentor(ctx, s1_def)
        M1_s2_def = n.Segmentor(ctx, s2_def)
        # the six Decoder call sites of the graph (dafnet.py:183-184,205-206 and the Z-regressor, :336-350) share
        # weights and the decoder has no batch statistics, so they run as ONE call on the concatenated batch:
        # y1 = Dec(s1,z1), y2 = Dec(s2,z2), y2_s1_def = Dec(s1_def,z2), y1_s2_def = Dec(s2_def,z1),
        # Dec(s1,z1_in), Dec(s2,z2_in)
        B = s1.shape[0]
        Z1_in, Z2_in = E.Var(z1_in), E.Var(z2_in)
        S_all = E.concat_rows(ctx, [s1, s2, s1_def, s2_def, s1, s2])
        Z_all = E.concat_rows(ctx, [z1, z2, z2, z1, Z1_in, Z2_in])
        y1, y2, y2_s1_def, y1_s2_def, yr1, yr2 = E.split_rows(ctx, n.Decoder(ctx, S_all, Z_all), [B] * 6)
        # Z-regressor branch
        z1_rec = n.Enc_Modality_mu(ctx, s1, yr1)
        z2_rec = n.Enc_Modality_mu(ctx, s2, yr2)

        # ---- losses
        if self.supervised:
            seg = [(M1, m1), (M2, m2), (M1_s2_def, m1), (M2_s1_def, m2)]
        else:
            seg = [(M1, m1), (M1_s2_def, m1)]
        for i, (pred, tgt) in enumerate(seg):
            E.loss_seg(ctx, pred, tgt, nm, True, c.w_sup_M, book.slot(i))
        for i, m in enumerate((M1, M2, M1_s2_def, M2_s1_def)):
            adv = n.D_Mask(ctx, E.slice_channels(ctx, m, 0, c.num_masks))
            E.loss_l1l2(ctx, adv, None, 1, c.w_adv_M, book.slot(ns + i), cval=1.0)
        for i, (y, tgt) in enumerate(((y1, x1), (y2, x2), (y1_s2_def, x1), (y2_s1_def, x2))):
            E.loss_l1l2(ctx, y, tgt, 0, c.w_rec_X, book.slot(ns + 4 + i))
        for i, (y, D) in enumerate(((y1, n.D_Image1), (y2, n.D_Image2), (y1_s2_def, n.D_Image1), (y2_s1_def, n.D_Image2))):
            E.loss_l1l2(ctx, D(ctx, y), None, 1, c.w_adv_X, book.slot(ns + 8 + i), cval=1.0)
        E.loss_l1l2(ctx, z1_rec, z1_in, 0, c.w_rec_Z, book.slot(ns + 14))
        E.loss_l1l2(ctx, z2_rec, z2_in, 0, c.w_rec_Z, book.slot(ns + 15))

    def pack(self, di, dt):
        x1, x2, z1, z2 = di[:4]
        B = x1.shape[0]
        eps1, eps2 = to_eps(B, self.net.conf.num_z), to_eps(B, self.net.conf.num_z)
        if self.supervised:
            return [x1, x2, z1, z2, eps1, eps2, dt[0], dt[1]]
        return [x1, x2, z1, z2, eps1, eps2, dt[0]]


class DAFNetPairedGeneratorTrainer(Trainer):
    """get_params_automated_pairing (dafnet.py:250-334) + the loss dict / weights of build_trainers_automatedpairs
    (dafnet.py:229-235).  Every modality brings n_pairs candidate images (candidate 0 = the expert pair); the candidates'
    deformed anatomies are decoded / segmented and their per-sample losses are mixed with the Balancer's weights.

    graph inputs : x1_1..x1_P, x2_1..x2_P, z1_in, z2_in, eps1, eps2, m1[, m2]
    outputs/loss : Segmentor x2 (x1): dice + .01*wBCE | SegmentorDef x2 (x1): sum_j w_j * per-sample(dice + .01*wBCE), w_sup_M
                   D_Mask x4: mse vs 1, w_adv_M | Decoder x2: mae, DecoderDef x2: sum_j w_j * per-sample mae, w_rec_X
                   D_Image1/2 x4: mse vs 1, w_adv_X | Enc_Modality x2: mean(KL), w_kl | ZReconstruct x2: mae, w_rec_Z
    """

    def __init__(self, net, supervised):
        c = net.conf
        self.net, self.supervised = net, supervised
        self.P = int(c.n_pairs)
        assert self.P == 3, "the Balancer takes 1 + 3 anatomies (model_components/balancer.py:17-20): n_pairs must be 3"
        seg = [("Segmentor", c.w_sup_M)] * 2 + [("SegmentorDef", c.w_sup_M)] * 2 if supervised else \
            [("Segmentor", c.w_sup_M), ("SegmentorDef", c.w_sup_M)]
        self.nseg = len(seg)
        nw = seg + [("D_Mask", c.w_adv_M)] * 4 + [("Decoder", c.w_rec_X)] * 2 + [("DecoderDef", c.w_rec_X)] * 2 + \
            [("D_Image1", c.w_adv_X), ("D_Image2", c.w_adv_X), ("D_Image1", c.w_adv_X), ("D_Image2", c.w_adv_X)] + \
            [("Enc_Modality", c.w_kl)] * 2 + [("ZReconstruct", c.w_rec_Z)] * 2
        super().__init__("supervised_trainer" if supervised else "unsupervised_trainer", net.generator_params(), c.lr, nw,
                         frozen_models=[net.D_Mask, net.D_Image1, net.D_Image2])

    def graph(self, ctx, book, *args):
        n, c, P = self.net, self.net.conf, self.P
        ns = self.nseg
        nm = n.loader.num_masks
        x1_lst, x2_lst = list(args[:P]), list(args[P:2 * P])
        rest = args[2 * P:]
        z1_in, z2_in, eps1, eps2, m1 = rest[:5]
        m2 = rest[5] if self.supervised else None
        x1, x2 = x1_lst[0], x2_lst[0]
        # encode: every candidate is a separate application of the encoder (its BatchNorm layers see one batch at a time)
        s1_lst = [n.Encoders_Anatomy[0](ctx, E.Var(x)) for x in x1_lst]
        s2_lst = [n.Encoders_Anatomy[1](ctx, E.Var(x)) for x in x2_lst]
        s1, s2 = s1_lst[0], s2_lst[0]
        X1, X2 = E.Var(x1), E.Var(x2)
        mu1, lv1 = n.Enc_Modality(ctx, s1, X1)
        mu2, lv2 = n.Enc_Modality(ctx, s2, X2)
        z1, _ = E.vae_sample(ctx, mu1, lv1, eps1, c.w_kl, book.slot(ns + 12))
        z2, _ = E.vae_sample(ctx, mu2, lv2, eps2, c.w_kl, book.slot(ns + 13))
        M1 = n.Segmentor(ctx, s1)
        M2 = n.Segmentor(ctx, s2)
        # deform every candidate towards the other modality's anatomy and weigh the candidates
        s1_def_lst = [n.Anatomy_Fuser.forward_deform(ctx, s1_i, s2) for s1_i in s1_lst]
        w1 = n.Balancer(ctx, s2, *s1_def_lst)
        s2_def_lst = [n.Anatomy_Fuser.forward_deform(ctx, s2_i, s1) for s2_i in s2_lst]
        w2 = n.Balancer(ctx, s1, *s2_def_lst)
        # all 4 + 2P Decoder call sites as one call (shared weights, no batch statistics)
        B = s1.shape[0]
        Z1_in, Z2_in = E.Var(z1_in), E.Var(z2_in)
        S_all = E.concat_rows(ctx, [s1, s2] + s1_def_lst + s2_def_lst + [s1, s2])
        Z_all = E.concat_rows(ctx, [z1, z2] + [z2] * P + [z1] * P + [Z1_in, Z2_in])
        ys = E.split_rows(ctx, n.Decoder(ctx, S_all, Z_all), [B] * (4 + 2 * P))
        y1, y2 = ys[0], ys[1]
        y2_s1_def_lst, y1_s2_def_lst = ys[2:2 + P], ys[2 + P:2 + 2 * P]
        yr1, yr2 = ys[2 + 2 * P], ys[3 + 2 * P]
        M1_s2_def_lst = [n.Segmentor(ctx, sd) for sd in s2_def_lst]
        M2_s1_def_lst = [n.Segmentor(ctx, sd) for sd in s1_def_lst]
        z1_rec = n.Enc_Modality_mu(ctx, s1, yr1)
        z2_rec = n.Enc_Modality_mu(ctx, s2, yr2)

        # ---- losses, in the order of all_outputs (dafnet.py:327-332)
        if self.supervised:
            E.loss_seg(ctx, M1, m1, nm, True, c.w_sup_M, book.slot(0))
            E.loss_seg(ctx, M2, m2, nm, True, c.w_sup_M, book.slot(1))
            E.loss_pairs(ctx, "seg", M1_s2_def_lst, m1, w2, c.w_sup_M, book.slot(2), nch=nm)
            E.loss_pairs(ctx, "seg", M2_s1_def_lst, m2, w1, c.w_sup_M, book.slot(3), nch=nm)
        else:
            E.loss_seg(ctx, M1, m1, nm, True, c.w_sup_M, book.slot(0))
            E.loss_pairs(ctx, "seg", M1_s2_def_lst, m1, w2, c.w_sup_M, book.slot(1), nch=nm)
        for i, m in enumerate((M1, M2, M1_s2_def_lst[0], M2_s1_def_lst[0])):
            adv = n.D_Mask(ctx, E.slice_channels(ctx, m, 0, c.num_masks))
            E.loss_l1l2(ctx, adv, None, 1, c.w_adv_M, book.slot(ns + i), cval=1.0)
        E.loss_l1l2(ctx, y1, x1, 0, c.w_rec_X, book.slot(ns + 4))
        E.loss_l1l2(ctx, y2, x2, 0, c.w_rec_X, book.slot(ns + 5))
        E.loss_pairs(ctx, "mae", y1_s2_def_lst, x1, w2, c.w_rec_X, book.slot(ns + 6))
        E.loss_pairs(ctx, "mae", y2_s1_def_lst, x2, w1, c.w_rec_X, book.slot(ns + 7))
        for i, (y, D) in enumerate(((y1, n.D_Image1), (y2, n.D_Image2), (y1_s2_def_lst[0], n.D_Image1),
                                    (y2_s1_def_lst[0], n.D_Image2))):
            E.loss_l1l2(ctx, D(ctx, y), None, 1, c.w_adv_X, book.slot(ns + 8 + i), cval=1.0)
        E.loss_l1l2(ctx, z1_rec, z1_in, 0, c.w_rec_Z, book.slot(ns + 14))
        E.loss_l1l2(ctx, z2_rec, z2_in, 0, c.w_rec_Z, book.slot(ns + 15))

    def pack(self, di, dt):
        """fit(x1_list + x2_list + [m1(, m2), z1, z2], targets) (dafnet_executor.py:447-454,470-477): the masks arrive
        among the INPUTS here"""
        P = self.P
        xs = list(di[:2 * P])
        B = xs[0].shape[0]
        eps1, eps2 = to_eps(B, self.net.conf.num_z), to_eps(B, self.net.conf.num_z)
        if self.supervised:
            m1, m2, z1, z2 = di[2 * P:2 * P + 4]
            return xs + [z1, z2, eps1, eps2, m1, m2]
        m1, z1, z2 = di[2 * P:2 * P + 3]
        return xs + [z1, z2, eps1, eps2, m1]


class _ZRegressorDAFNet(Trainer):
    """dafnet.py:336-350 (compiled with its own Adam; in DAFNet only used nested inside the trainers)"""

    def __init__(self, net):
        c = net.conf
        self.net = net
        params = net.Decoder.params() + [p for l in net.Enc_Modality.mu_layers for p in l.params()]
        super().__init__("ZReconstruct", params, c.lr, [("ZReconstruct", c.w_rec_Z)] * 2)

    def graph(self, ctx, book, s1, s2, z1, z2):
        for i, (s, z) in enumerate(((s1, z1), (s2, z2))):
            E.loss_l1l2(ctx, self.net.z_reconstruct(ctx, E.Var(s), E.Var(z)), z, 0, self.net.conf.w_rec_Z, book.slot(i))

    def pack(self, di, dt):
        return list(di)


class DAFNet(MMSDNet):
    def __init__(self, conf):
        super(DAFNet, self).__init__(conf)
        self.D_Image1 = None
        self.D_Image2 = None
        self.Balancer = None
        self.D_Image1_trainer = None
        self.D_Image2_trainer = None

    def build(self):
        self.build_mask_discriminator()
        self.build_image_discriminator1()
        self.build_image_discriminator2()
        self.build_generators()
        try:
            self.load_models()
        except Exception:
            log.warning("No models found")

    def _components(self):
        return [("D_Mask", self.D_Mask), ("D_Image1", self.D_Image1), ("D_Image2", self.D_Image2),
                ("Enc_Anatomy1", self.Encoders_Anatomy[0]), ("Enc_Anatomy2", self.Encoders_Anatomy[1]),
                ("Enc_Modality", self.Enc_Modality), ("Anatomy_Fuser", self.Anatomy_Fuser),
                ("Segmentor", self.Segmentor), ("Decoder", self.Decoder)]

    def load_models(self):
        """dafnet.py:54-73 -- raises when the files are absent (caught by build())"""
        model_folder = self.conf.folder + "/models/"
        for fname, m in self._components():
            m.load_weights(model_folder + fname)
        try:
            self.Balancer.load_weights(model_folder + "Balancer")
        except Exception:
            pass
        log.info("Loading trained models from file")

    def save_models(self):
        model_folder = self.conf.folder + "/models/"
        for fname, m in self._components() + [("Balancer", self.Balancer)]:
            m.save_weights(model_folder + fname)

    def _build_image_discriminator(self, name):
        params = self.conf.d_image_params
        params["name"] = name           # the reference mutates one shared dict (dafnet.py:79-80,100-101)
        with BuildScope(rng=self.rng) as sc:
            D = Discriminator(params)
            D.build()
        sc.arena.to_device()
        sc.state.to_device()
        D.model.summary(print_fn=log.info)
        return D.model, DiscriminatorTrainer(name + "_trainer", D.model, self.conf.d_image_params.lr)

    def build_image_discriminator1(self):
        self.D_Image1, self.D_Image1_trainer = self._build_image_discriminator("D_Image1")

    def build_image_discriminator2(self):
        self.D_Image2, self.D_Image2_trainer = self._build_image_discriminator("D_Image2")

    def build_generators(self):
        assert self.D_Mask is not None, "Discriminator has not been built yet"
        # frozen inside the generator trainers (dafnet.py:119-121); the D trainers re-enable their own weights
        with self.gen_scope:
            self.Encoders_Anatomy = AnatomyEncoders(self.modalities).build(self.conf.anatomy_encoder)
            self.Anatomy_Fuser = anatomy_fuser.build(self.conf)
            self.Enc_Modality = modality_encoder.build(self.conf)
            self.Enc_Modality_mu = MuModel(self.Enc_Modality)
            self.Segmentor = segmentor.build(self.conf)
            self.Decoder = decoder.build(self.conf)
            self.Balancer = balancer.build(self.conf)
        self.gen_scope.arena.to_device()
        self.gen_scope.state.to_device()
        self.build_trainers()

    def build_trainers(self):
        self.build_z_regressor()
        if not self.conf.automatedpairing:
            self.build_trainers_expertpairs()
        else:
            self.build_trainers_automatedpairs()

    def build_trainers_expertpairs(self):
        self.unsupervised_trainer = DAFNetGeneratorTrainer(self, supervised=False)
        self.supervised_trainer = DAFNetGeneratorTrainer(self, supervised=True)

    def build_trainers_automatedpairs(self):
        """dafnet.py:224-248"""
        self.unsupervised_trainer = DAFNetPairedGeneratorTrainer(self, supervised=False)
        self.supervised_trainer = DAFNetPairedGeneratorTrainer(self, supervised=True)

    def generator_params(self):
        """the automated-pairing trainers also train the Balancer (it is nested in their graphs, dafnet.py:283-287)"""
        out = super(DAFNet, self).generator_params()
        if getattr(self.conf, "automatedpairing", False):
            out = out + [p for p in self.Balancer.params() if all(p is not q for q in out)]
        return out

    def build_z_regressor(self):
        self.Z_Regressor = _ZRegressorDAFNet(self)

    def calculate_weights(self, inputs):
        """dafnet.py:352-361 (inference form): Balancer weights for [s_mod2] + s_list"""
        if len(inputs[1:]) == 1:
            return None
        return self.Balancer.predict_device(*inputs)

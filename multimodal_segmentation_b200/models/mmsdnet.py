"""MMSDNet (reference: models/mmsdnet.py:17-232): two independent UNet anatomy encoders, modality
encoder, segmentor, decoder, anatomy fuser and a mask discriminator, wired into supervised /
unsupervised trainers and a separately optimised Z-regressor."""
import logging
import os

import numpy as np

from .. import engine as E
from ..keras_like import BuildScope, Model
from ..model_components import anatomy_encoder, anatomy_fuser, decoder, modality_encoder, segmentor
from .basenet import BaseNet
from .discriminator import Discriminator
from .trainers import DiscriminatorTrainer, Trainer

log = logging.getLogger("mmsdnet")


def MuModel(enc):
    """Enc_Modality_mu = Model(Enc_Modality.inputs, Enc_Modality.get_layer('z_mean').output)   (models/mmsdnet.py:51,87;
    models/dafnet.py:126): a sub-model that shares the encoder's layers up to ``z_mean``"""
    return Model(enc.inputs, enc.get_layer("z_mean").output, name="Enc_Modality_mu")


class MMSDNetGeneratorTrainer(Trainer):
    """models/mmsdnet.py:95-192.  graph inputs: x1, x2, eps[6] ; targets as fed by
    model_executors/mmsdnet_executor.py:257-260 (supervised) / :287-290 (unsupervised)."""

    def __init__(self, net, supervised):
        c = net.conf
        self.net, self.supervised = net, supervised
        nseg = 6 if supervised else 3
        nw = [("Segmentor", c.w_sup_M)] * nseg + [("D_Mask", c.w_adv_M)] * 6 + [("Decoder", c.w_rec_X)] * 6 + \
             [("Enc_Modality", c.w_kl)] * 6
        super().__init__("supervised_trainer" if supervised else "unsupervised_trainer", net.generator_params(), c.lr, nw,
                         frozen_models=[net.D_Mask])

    def graph(self, ctx, book, x1, x2, eps, seg_targets, rec_targets):
        n, c = self.net, self.net.conf
        x = [E.Var(x1), E.Var(x2)]
        nseg = 6 if self.supervised else 3
        kslot = nseg + 12
        s = [n.Encoders_Anatomy[i](ctx, x[i]) for i in range(2)]
        z = []
        for i in range(2):
            mu, lv = n.Enc_Modality(ctx, s[i], x[i])
            z.append(E.vae_sample(ctx, mu, lv, eps[i], c.w_kl, book.slot(kslot + i))[0])
        m1, m2 = n.Segmentor(ctx, s[0]), n.Segmentor(ctx, s[1])
        rec = [n.Decoder(ctx, s[i], z[i]) for i in range(2)]
        s1_def, s1_fused = n.Anatomy_Fuser(ctx, s[0], s[1])
        s2_def, s2_fused = n.Anatomy_Fuser(ctx, s[1], s[0])
        fused_seg = [n.Segmentor(ctx, a) for a in (s1_def, s1_fused, s2_def, s2_fused)]
        m_list = ([m1, m2] + fused_seg) if self.supervised else ([m1] + fused_seg[2:])
        adv_in = [m1, m2] + fused_seg
        zl1 = []
        for j, a in enumerate((s1_def, s1_fused)):
            mu, lv = n.Enc_Modality(ctx, a, x[1])
            zl1.append(E.vae_sample(ctx, mu, lv, eps[2 + j], c.w_kl, book.slot(kslot + 2 + j))[0])
        rec += [n.Decoder(ctx, a, zl1[j]) for j, a in enumerate((s1_def, s1_fused))]
        zl2 = []
        for j, a in enumerate((s2_def, s2_fused)):
            mu, lv = n.Enc_Modality(ctx, a, x[0])
            zl2.append(E.vae_sample(ctx, mu, lv, eps[4 + j], c.w_kl, book.slot(kslot + 4 + j))[0])
        rec += [n.Decoder(ctx, a, zl2[j]) for j, a in enumerate((s2_def, s2_fused))]
        nm = n.loader.num_masks
        for i, (m, tgt) in enumerate(zip(m_list, seg_targets)):
            E.loss_seg(ctx, m, tgt, nm, False, c.w_sup_M, book.slot(i))      # dice only (mmsdnet.py:184)
        for i, m in enumerate(adv_in):
            adv = n.D_Mask(ctx, E.slice_channels(ctx, m, 0, c.num_masks))
            E.loss_l1l2(ctx, adv, None, 1, c.w_adv_M, book.slot(nseg + i), cval=1.0)
        for i, (y, tgt) in enumerate(zip(rec, rec_targets)):
            E.loss_l1l2(ctx, y, tgt, 0, c.w_rec_X, book.slot(nseg + 6 + i))

    def pack(self, di, dt):
        x1, x2 = di[:2]
        nseg = 6 if self.supervised else 3
        B = x1.shape[0]
        eps = [to_eps(B, self.net.conf.num_z) for _ in range(6)]
        return [x1, x2, eps, dt[:nseg], dt[nseg + 6:nseg + 12]]


def to_eps(B, Z):
    import torch
    return torch.from_numpy(np.random.normal(0, 1, (B, Z)).astype(np.float32)).cuda()


class ZRegressorTrainer(Trainer):
    """models/mmsdnet.py:194-208 / models/dafnet.py:336-350: y = Decoder([s, z]); z_rec = Enc_Modality_mu([s, y]);
    loss mae(z, z_rec) * w_rec_Z; its own Adam over the Decoder + Enc_Modality(mu path) weights."""

    def __init__(self, net, num_inputs):
        c = net.conf
        self.net, self.num_inputs = net, num_inputs
        params = net.Decoder.params() + [p for l in net.Enc_Modality.mu_layers for p in l.params()]
        super().__init__("ZReconstruct", params, c.lr, [("ZReconstruct", c.w_rec_Z)] * num_inputs)

    def graph(self, ctx, book, *args):
        k = self.num_inputs
        s_list, z_list = args[:k], args[k:2 * k]
        for i in range(k):
            z_rec = self.net.z_reconstruct(ctx, E.Var(s_list[i]), E.Var(z_list[i]))
            E.loss_l1l2(ctx, z_rec, z_list[i], 0, self.net.conf.w_rec_Z, book.slot(i))

    def pack(self, di, dt):
        return list(di)


class MMSDNet(BaseNet):
    def __init__(self, conf):
        super(MMSDNet, self).__init__(conf)
        self.modalities = conf.modality
        self.D_Mask = None
        self.Encoders_Anatomy = None
        self.Enc_Modality = None
        self.Enc_Modality_mu = None
        self.Anatomy_Fuser = None
        self.Segmentor = None
        self.Decoder = None
        self.D_Mask_trainer = None
        self.unsupervised_trainer = None
        self.supervised_trainer = None
        self.Z_Regressor = None
        seed = getattr(conf, "seed", 0)
        self.rng = np.random.RandomState(seed)
        self.gen_scope = BuildScope(rng=self.rng)

    # ------------------------------------------------------------------ build
    def build(self):
        self.build_mask_discriminator()
        self.build_generators()
        self.load_models()

    def generator_params(self):
        comps = list(self.Encoders_Anatomy) + [self.Enc_Modality, self.Anatomy_Fuser, self.Segmentor, self.Decoder]
        out, seen = [], set()
        for m in comps:
            for p in m.params():
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out

    def z_reconstruct(self, ctx, s, z):
        y = self.Decoder(ctx, s, z)
        return self.Enc_Modality_mu(ctx, s, y)

    def load_models(self):
        f = self.conf.folder + "/supervised_trainer.npz"
        if os.path.exists(f):
            log.info("Loading trained models from file")
            z = np.load(f)
            for m in list(self.Encoders_Anatomy) + [self.Enc_Modality, self.Anatomy_Fuser, self.Segmentor, self.Decoder,
                                                    self.D_Mask]:
                m.set_weights([z[m.name + "/" + p.name] for p in m.weight_list()])

    def save_models(self):
        log.debug("Saving trained models")
        os.makedirs(self.conf.folder, exist_ok=True)
        d = {}
        for m in list(self.Encoders_Anatomy) + [self.Enc_Modality, self.Anatomy_Fuser, self.Segmentor, self.Decoder,
                                                self.D_Mask]:
            for p, w in zip(m.weight_list(), m.get_weights()):
                d[m.name + "/" + p.name] = w
        np.savez(self.conf.folder + "/supervised_trainer.npz", **d)

    def build_mask_discriminator(self):
        with BuildScope(rng=self.rng) as sc:
            D = Discriminator(self.conf.d_mask_params)
            D.build()
        sc.arena.to_device()
        sc.state.to_device()
        log.info("Mask Discriminator D_M")
        D.model.summary(print_fn=log.info)
        self.D_Mask = D.model
        self.D_Mask_trainer = DiscriminatorTrainer("D_Mask_trainer", self.D_Mask, self.conf.d_mask_params.lr)

    def build_generators(self):
        assert self.D_Mask is not None, "Discriminator has not been built yet"
        with self.gen_scope:
            self.Encoders_Anatomy = [anatomy_encoder.build(self.conf.anatomy_encoder, "Enc_Anatomy_%s" % mod)
                                     for mod in self.modalities]
            # distinct weight names for the two independent UNets
            for i, m in enumerate(self.Encoders_Anatomy):
                for p in m.weight_list():
                    p.name = "enc%d_%s" % (i + 1, p.name)
            self.Anatomy_Fuser = anatomy_fuser.build(self.conf)
            self.Enc_Modality = modality_encoder.build(self.conf)
            self.Enc_Modality_mu = MuModel(self.Enc_Modality)
            self.Segmentor = segmentor.build(self.conf)
            self.Decoder = decoder.build(self.conf)
        self.gen_scope.arena.to_device()
        self.gen_scope.state.to_device()
        self.build_unsupervised_trainer()
        self.build_supervised_trainer()
        self.build_z_regressor()

    def build_unsupervised_trainer(self):
        self.unsupervised_trainer = MMSDNetGeneratorTrainer(self, supervised=False)

    def build_supervised_trainer(self):
        self.supervised_trainer = MMSDNetGeneratorTrainer(self, supervised=True)

    def build_z_regressor(self):
        self.Z_Regressor = ZRegressorTrainer(self, len(self.modalities) + 4)

    # ------------------------------------------------------------------ inference
    def predict_mask(self, modality_index, type, image_list):
        """models/mmsdnet.py:210-232"""
        assert type in ["simple", "def", "max", "maxnostn"]
        idx2 = modality_index
        idx1 = 1 - idx2
        images_mod1 = image_list[idx1]
        images_mod2 = image_list[idx2]
        s1 = self.Encoders_Anatomy[idx1].predict(images_mod1)
        s2 = self.Encoders_Anatomy[idx2].predict(images_mod2)
        if type == "simple":
            return self.Segmentor.predict(s2)
        elif type == "def":
            return self.Segmentor.predict(self.Anatomy_Fuser.predict([s1, s2])[0])
        elif type == "max":
            return self.Segmentor.predict(self.Anatomy_Fuser.predict([s1, s2])[1])
        elif type == "maxnostn":
            return self.Segmentor.predict(np.max([s1, s2], axis=0))
        raise ValueError(type)

    def predict_mask_device(self, modality_index, type, x_a, x_b):
        """same as predict_mask with device tensors end to end (no host round trips)"""
        idx2 = modality_index
        idx1 = 1 - idx2
        xs = [x_a, x_b]
        s2 = self.Encoders_Anatomy[idx2].predict_device(xs[idx2])
        if type == "simple":
            return self.Segmentor.predict_device(s2)
        s1 = self.Encoders_Anatomy[idx1].predict_device(xs[idx1])
        d, f = self.Anatomy_Fuser.predict_device(s1, s2)
        return self.Segmentor.predict_device(d if type == "def" else f)

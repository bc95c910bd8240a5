"""MMSDNet executor (reference: model_executors/mmsdnet_executor.py:159-331).

train_batch = train_batch_generators (one Adam step of the supervised / unsupervised trainer, then one Adam step of the
Z-regressor on anatomies predicted in inference mode, :242-306) + train_batch_mask_discriminator (:308-331).
Differences from DAFNet that are kept: the fused (Maximum) anatomy IS trained (models/mmsdnet.py:160-165), no image
discriminators, masks are fed WITHOUT the residual channel (dice slices the 5-channel prediction, costs.py:62-65),
the Z-regressor is its own optimizer step.  Host -> device staging is separated from the device work exactly as
in the DAFNet executor."""
import logging

import numpy as np
import torch

from .. import costs, ops
from .dafnet_executor import DAFNetExecutor

log = logging.getLogger("mmsdnet_executor")


class MMSDNetExecutor(DAFNetExecutor):
    # mmsdnet_executor.py:159-236: no stochastic weight averaging -- validation (and with it the early-stopping metric
    # val_loss_mod2_fused) runs on the LIVE models and model.save_models() writes the live weights
    USE_SWA = False

    def get_loss_names(self):
        return ["adv_M", "rec_X", "dis_M", "val_loss", "val_loss_mod1", "val_loss_mod2", "val_loss_mod2_s1def",
                "val_loss_mod2_fused", "supervised_Mask", "loss", "KL", "rec_Z"]

    # ------------------------------------------------------------------ data (mmsdnet_executor.py:67-157)
    def _init_labelled_data_generator(self):
        if self.conf.l_mix == 0:
            return None
        self.data = self.loader.load_all_modalities_concatenated(self.conf.split, "training", self.conf.image_downsample)
        self.data.sample(int(np.round(self.conf.l_mix * self.data.num_volumes)), seed=self.conf.seed)
        self.data_len = self.data.size()
        nm = self.loader.num_masks
        masks = [np.ascontiguousarray(self.data.get_masks_modi(i)[..., 0:nm]) for i in range(2)]      # no residual channel
        return self.get_data_generator(train_images=[self.data.get_images_modi(i) for i in range(2)], train_labels=masks)

    def _init_unlabelled_data_generator(self):
        if self.conf.l_mix == 1:
            return None
        self.ul_data = self._load_unlabelled_data("training")        # the volumes the labelled sample left out
        if self.data is None or self.ul_data.size() > self.data.size():
            self.data_len = self.ul_data.size()
        nm = self.loader.num_masks
        return self.get_data_generator(train_images=[self.ul_data.get_images_modi(i) for i in range(2)],
                                       train_labels=[np.ascontiguousarray(self.ul_data.get_masks_modi(0)[..., 0:nm])])

    def validate(self, epoch_loss):
        """mmsdnet_executor.py:210-236: 1 - Dice(binarised) on the validation split through the LIVE models: modality 1,
        modality 2 from its own anatomy, from the deformed modality-1 anatomy ('s1def') and from the fused anatomy"""
        valid = self.loader.load_all_modalities_concatenated(self.conf.split, "validation", self.conf.image_downsample)
        valid.crop(self.conf.input_shape[:2])
        nm = self.loader.num_masks
        x0, x1 = valid.get_images_modi(0), valid.get_images_modi(1)
        real0, real1 = valid.get_masks_modi(0)[..., :nm], valid.get_masks_modi(1)[..., :nm]
        M = self.model
        s1 = M.Encoders_Anatomy[0].predict(x0)
        s2 = M.Encoders_Anatomy[1].predict(x1)
        s1_deformed, s_fused = M.Anatomy_Fuser.predict([s1, s2])
        loss = lambda real, s: 1 - costs.dice(real, M.Segmentor.predict(s), binarise=True)
        l_mod1, l_mod2 = loss(real0, s1), loss(real1, s2)
        l_mod2_s1def, l_mod2_fused = loss(real1, s1_deformed), loss(real1, s_fused)
        epoch_loss["val_loss_mod2"].append(l_mod2)
        epoch_loss["val_loss_mod2_s1def"].append(l_mod2_s1def)
        epoch_loss["val_loss_mod2_fused"].append(l_mod2_fused)
        epoch_loss["val_loss_mod1"].append(l_mod1)
        epoch_loss["val_loss"].append(np.mean([l_mod1, l_mod2, l_mod2_s1def, l_mod2_fused]))

    # ------------------------------------------------------------------ staging
    def _stage_generator(self, supervised):
        batch = self._stage(self.gen_labelled if supervised else self.gen_unlabelled)
        B = batch[0].shape[0]
        # 6 reparameterisation noises (one per Enc_Modality call site) + 6 sampled z for the Z-regressor
        return batch + [self._sample_z(B) for _ in range(12)]

    def _stage_mask_d(self):
        (m,) = self._stage(self.discriminator_masks)
        (x1,) = self._stage(self.discriminator_image[0])
        (x2,) = self._stage(self.discriminator_image[1])
        B = min(t.shape[0] for t in (x1, x2, m))
        nm = self.conf.num_masks
        return [x1[:B].contiguous(), x2[:B].contiguous(), m[:B, ..., 0:nm].contiguous(), self._sample_idx(4 * B, B)]

    def stage_step_inputs(self):
        """mmsdnet_executor.py:238-240: train_batch = the generator updates of the labelled and / or the unlabelled
        branch (each followed by its Z-regressor update), then ONE mask-discriminator update whatever l_mix.  Step
        entries are (kind, generator inputs or None, mask-discriminator inputs or None, [])."""
        step = []
        if self.conf.l_mix > 0:
            step.append(("sup", self._stage_generator(True), None, []))
        if self.conf.l_mix < 1:
            step.append(("unsup", self._stage_generator(False), None, []))
        step.append(("dmask", None, self._stage_mask_d(), []))
        return step

    def _run_step(self, step):
        for kind, g, dm, _ in step:
            if g is not None:
                self._run_generator(kind == "sup", g)
            if dm is not None:
                self._run_mask_d(dm)

    # ------------------------------------------------------------------ step (mmsdnet_executor.py:238-331)
    def train_batch(self, epoch_loss):
        if self._graph is not None:
            return super(MMSDNetExecutor, self).train_batch(epoch_loss)
        self._run_step(self.stage_step_inputs())

    def train_batch_on(self, step):
        if self._graph is not None:
            return self.train_batch_graph(step)
        self._run_step(step)

    def train_batch_generators(self, epoch_loss):
        if self.conf.l_mix > 0:
            self._run_generator(True, self._stage_generator(True))
        if self.conf.l_mix < 1:
            self._run_generator(False, self._stage_generator(False))

    def train_batch_mask_discriminator(self, epoch_loss):
        self._run_mask_d(self._stage_mask_d())

    def _run_image_d(self, t):
        return None

    def _run_generator(self, supervised, t):
        M = self.model
        if supervised:
            x1, x2, m1, m2 = t[:4]
            noise = t[4:]
            seg_targets = [m1, m2, m2, m2, m1, m1]            # mmsdnet_executor.py:257
            tr = M.supervised_trainer
        else:
            x1, x2, m1 = t[:3]
            noise = t[3:]
            seg_targets = [m1, m1, m1]                        # :287
            tr = M.unsupervised_trainer
        eps, z_list = list(noise[:6]), list(noise[6:12])
        rec_targets = [x1, x2, x2, x2, x1, x1]                # :259
        tr.train_on_device(x1, x2, eps, seg_targets, rec_targets)
        self._pending.append((tr, tr.book.snapshot(), "gen"))
        # Z regressor on anatomies predicted in inference mode (:266-275)
        s1 = M.Encoders_Anatomy[0].predict_device(x1)
        s2 = M.Encoders_Anatomy[1].predict_device(x2)
        s1_def, s1_fused = M.Anatomy_Fuser.predict_device(s1, s2)
        s2_def, s2_fused = M.Anatomy_Fuser.predict_device(s2, s1)
        M.Z_Regressor.train_on_device(s1, s2, s1_def, s1_fused, s2_def, s2_fused, *z_list)
        self._pending.append((M.Z_Regressor, M.Z_Regressor.book.snapshot(), "rec_Z"))

    def _run_mask_d(self, t):
        x1, x2, m, idx = t
        M = self.model
        nm = self.conf.num_masks
        B = x1.shape[0]
        fake_s = [M.Encoders_Anatomy[0].predict_device(x1), M.Encoders_Anatomy[1].predict_device(x2)]
        fake_m = [M.Segmentor.predict_device(s) for s in fake_s]
        s1_def, s1_fused = M.Anatomy_Fuser.predict_device(fake_s[0], fake_s[1])
        fake_m += [M.Segmentor.predict_device(s) for s in (s1_def, s1_fused)]
        cat = torch.empty((4 * B,) + tuple(fake_m[0].shape[1:3]) + (nm,), dtype=torch.float32, device="cuda")
        for i, fm in enumerate(fake_m):
            ops.copy_channels(fm, 0, cat[i * B:(i + 1) * B], 0, nm)
        fake = ops.gather_rows(cat, idx)
        M.D_Mask_trainer.train_on_device(m, fake)
        self._pending.append((M.D_Mask_trainer, M.D_Mask_trainer.book.snapshot(), "dis_M"))

    def flush_losses(self, epoch_loss):
        for tr, snap, kind in self._pending:
            h = tr.book.history(snap)
            self.d2h_bytes += snap.numel() * 4
            if kind == "gen":
                epoch_loss["supervised_Mask"].append(h["Segmentor_loss"][0])
                epoch_loss["adv_M"].append(h["D_Mask_loss"][0])
                epoch_loss["rec_X"].append(h["Decoder_loss"][0])
                epoch_loss["KL"].append(h["Enc_Modality_loss"][0])
                epoch_loss["loss"].append(h["loss"][0])
            else:
                epoch_loss[kind].append(h["loss"][0])
        self._pending = []

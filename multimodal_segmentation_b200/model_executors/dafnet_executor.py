"""DAFNet executor (reference: model_executors/dafnet_executor.py:21-583).

``train_batch`` keeps the reference's step schedule (dafnet_executor.py:369-387):
  generator update (supervised and/or unsupervised trainer) -> two mask-discriminator updates ->
  image-discriminator-1 and -2 updates, with the fake samples generated in the inference phase on
  freshly drawn batches and randomly sub-sampled (utils/data_utils.py:125-129).
The reference makes 25 session.run round trips per train_batch (numpy out, numpy in); here every
intermediate stays in HBM: the only host<->device traffic of a step is the H2D copy of the input
batches from pinned memory and the D2H read of the loss slots.
"""
import logging
import os

import numpy as np
import torch

from .. import costs
from .. import ops
from ..callbacks.swa import SWA
from ..model_components import anatomy_encoder, anatomy_fuser, balancer, decoder, modality_encoder, segmentor
from ..models.discriminator import Discriminator
from ..utils import data_utils
from ..utils.distributions import NormalDistribution
from .base_executor import Executor

log = logging.getLogger("dafnet_executor")


def _h2d(t):
    return t.cuda(non_blocking=True) if torch.is_tensor(t) else torch.from_numpy(np.ascontiguousarray(t, np.float32)).cuda()


class DAFNetExecutor(Executor):
    def __init__(self, conf, model):
        super(DAFNetExecutor, self).__init__(conf, model)
        self.loader.modalities = self.conf.modality
        self.gen_labelled = None
        self.gen_unlabelled = None
        self.discriminator_masks = None
        self.discriminator_image = None
        self.data = None
        self.ul_data = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._pending = []
        self._graph = None
        self._static = None
        self._graph_pending = []
        self._next = None
        self._next_ready = None      # event: the prefetched step's copies have finished on the copy stream
        self._copy_stream = None
        self.init_swa_models()

    # ------------------------------------------------------------------ stochastic weight averaging
    SWA_EPOCH = 40
    USE_SWA = True          # the reference's MMSDNet executor has no SWA (mmsdnet_executor.py:159-208): subclass sets False

    def init_swa_models(self):
        """dafnet_executor.py:41-56"""
        if not self.USE_SWA:
            return
        e = self.SWA_EPOCH
        self.swa_D_Mask = SWA(e, Discriminator(self.conf.d_mask_params).build, None)
        has_dimg = hasattr(self.conf, "d_image_params")
        self.swa_D_Image1 = SWA(e, Discriminator(self.conf.d_image_params).build, None) if has_dimg else None
        self.swa_D_Image2 = SWA(e, Discriminator(self.conf.d_image_params).build, None) if has_dimg else None
        self.swa_Enc_Anatomy1 = SWA(e, anatomy_encoder.build, self.conf.anatomy_encoder)
        self.swa_Enc_Anatomy2 = SWA(e, anatomy_encoder.build, self.conf.anatomy_encoder)
        self.swa_Enc_Modality = SWA(e, modality_encoder.build, self.conf)
        self.swa_Anatomy_Fuser = SWA(e, anatomy_fuser.build, self.conf)
        self.swa_Segmentor = SWA(e, segmentor.build, self.conf)
        self.swa_Decoder = SWA(e, decoder.build, self.conf)
        self.swa_Balancer = SWA(e, balancer.build, self.conf) if getattr(self.model, "Balancer", None) is not None else None
        self.set_swa_model_weights()

    def set_swa_model_weights(self):
        """dafnet_executor.py:58-68"""
        if not self.USE_SWA:
            return
        M = self.model
        self.swa_D_Mask.model = M.D_Mask
        if self.swa_D_Image1 is not None:
            self.swa_D_Image1.model = getattr(M, "D_Image1", None)
            self.swa_D_Image2.model = getattr(M, "D_Image2", None)
        self.swa_Enc_Anatomy1.model = M.Encoders_Anatomy[0]
        self.swa_Enc_Anatomy2.model = M.Encoders_Anatomy[1]
        self.swa_Enc_Modality.model = M.Enc_Modality
        self.swa_Anatomy_Fuser.model = M.Anatomy_Fuser
        self.swa_Segmentor.model = M.Segmentor
        self.swa_Decoder.model = M.Decoder
        if self.swa_Balancer is not None:
            self.swa_Balancer.model = M.Balancer

    def get_swa_models(self):
        """dafnet_executor.py:207-210"""
        if not self.USE_SWA:
            return []
        lst = [self.swa_D_Mask, self.swa_D_Image1, self.swa_D_Image2, self.swa_Enc_Anatomy1, self.swa_Enc_Anatomy2,
               self.swa_Enc_Modality, self.swa_Anatomy_Fuser, self.swa_Segmentor, self.swa_Decoder, self.swa_Balancer]
        return [m for m in lst if m is not None and m.model is not None]

    def save_models(self, postfix=""):
        """dafnet_executor.py:286-301: the files hold the SWA weights (identical to the live ones until SWA_EPOCH)"""
        if not self.USE_SWA or not hasattr(self.model, "_components"):
            return self.model.save_models()         # MMSDNet: live weights, single-file format (models/mmsdnet.py:42-60)
        model_folder = self.conf.folder + "/models/"
        os.makedirs(model_folder, exist_ok=True)
        names = [("D_Mask", self.swa_D_Mask), ("D_Image1", self.swa_D_Image1), ("D_Image2", self.swa_D_Image2),
                 ("Enc_Anatomy1", self.swa_Enc_Anatomy1), ("Enc_Anatomy2", self.swa_Enc_Anatomy2),
                 ("Enc_Modality", self.swa_Enc_Modality), ("Anatomy_Fuser", self.swa_Anatomy_Fuser),
                 ("Segmentor", self.swa_Segmentor), ("Decoder", self.swa_Decoder), ("Balancer", self.swa_Balancer)]
        for fname, swa_m in names:
            if swa_m is None or swa_m.model is None:
                continue
            # same arrays a clone would hold (get_clone_model().save_weights in the reference), without building it
            ws = swa_m.swa_weights if swa_m.swa_weights is not None else swa_m.model.get_weights()
            order = [p.name for p in swa_m.model.weight_list()]
            np.savez(model_folder + fname + postfix + ".npz", __order__=np.array(order), **dict(zip(order, ws)))

    # ------------------------------------------------------------------ data
    def init_train_data(self):
        self.gen_labelled = self._init_labelled_data_generator()
        self.gen_unlabelled = self._init_unlabelled_data_generator()
        self.discriminator_masks = self._init_disciminator_mask_generator()
        self.discriminator_image = [self._init_discriminator_image_generator(mod) for mod in self.model.modalities]
        self.batches = int(np.ceil(self.data_len / self.conf.batch_size))

    def _init_labelled_data_generator(self):
        if self.conf.l_mix == 0:
            return None
        self.data = self.loader.load_all_modalities_concatenated(self.conf.split, "training", self.conf.image_downsample)
        self.data.sample(int(np.round(self.conf.l_mix * self.data.num_volumes)), seed=self.conf.seed)
        self._pair_up(self.data)
        self.data_len = self.data.size()
        nm = self.loader.num_masks
        # add_residual (dafnet_executor.py:493-494): the channel is allocated here; with augmentation the stager rebuilds
        # it from the ROTATED mask channels, as the reference does (the background is 1 wherever no channel is exactly 1)
        masks = [self.add_residual(self.data.get_masks_modi(i)[..., 0:nm]) for i in range(2)]
        return self.get_data_generator(train_images=[self.data.get_images_modi(i) for i in range(2)], train_labels=masks,
                                       labels_have_residual=True)

    def _init_unlabelled_data_generator(self):
        if self.conf.l_mix == 1:
            return None
        self.ul_data = self._load_unlabelled_data("training")
        if self.data is None or self.ul_data.size() > self.data.size():
            self.data_len = self.ul_data.size()
        nm = self.loader.num_masks
        return self.get_data_generator(train_images=[self.ul_data.get_images_modi(i) for i in range(2)],
                                       train_labels=[self.add_residual(self.ul_data.get_masks_modi(0)[..., 0:nm])],
                                       labels_have_residual=True)

    def _load_unlabelled_data(self, split_type):
        """dafnet_executor.py:117-145 ('ul'): the training volumes that are NOT labelled -- the same draw of
        round(l_mix * num_volumes) volumes as Data.sample made for the labelled set (same seed), removed"""
        ul_data = self.loader.load_all_modalities_concatenated(self.conf.split, split_type, self.conf.image_downsample)
        self._pair_up(ul_data)
        if self.conf.l_mix > 0:
            num_lb_vols = int(np.round(self.conf.l_mix * ul_data.num_volumes))
            labelled = set(np.asarray(ul_data.get_sample_volumes(num_lb_vols, seed=self.conf.seed)).tolist())
            ul_data.filter_volumes([v for v in ul_data.volume_ids() if v not in labelled])
        return ul_data

    def _pair_up(self, data):
        """dafnet_executor.py:89-93,127-131: candidate pairs for the automated-pairing trainers (images get n_pairs
        channels, channel 0 = expert pair) or randomised pairs"""
        if getattr(self.conf, "randomise", False):
            data.randomise_pairs(self.conf.n_pairs - 1, seed=self.conf.seed)
        elif getattr(self.conf, "automatedpairing", False):
            np.random.seed(self.conf.seed)
            data.expand_pairs(self.conf.n_pairs - 1, 0, neighborhood=self.conf.n_pairs)
            data.expand_pairs(self.conf.n_pairs - 1, 1, neighborhood=self.conf.n_pairs)

    def _init_disciminator_mask_generator(self):
        """real masks for D_Mask (dafnet_executor.py:147-176): both modalities' masks of the labelled volumes and modality
        1's masks of the unlabelled ones"""
        masks = []
        if self.data is not None:
            masks.append(np.concatenate([self.data.get_masks_modi(0), self.data.get_masks_modi(1)], axis=0))
        if self.ul_data is not None and self.ul_data.size() > 0:
            masks.append(self.ul_data.get_masks_modi(0))
        if masks:
            masks = np.concatenate(masks, axis=0)
        else:
            masks = np.empty((0,) + tuple(self.conf.input_shape[:-1]) + (self.loader.num_masks,), np.float32)
        assert masks.shape[1:3] == tuple(self.conf.input_shape[:2]), masks.shape
        return self.get_data_generator(train_images=None, train_labels=[masks])

    def _init_discriminator_image_generator(self, modality):
        """every training image of one modality, labelled or not (dafnet_executor.py:178-184, data_type 'all')"""
        d = self.loader.load_all_modalities_concatenated(self.conf.split, "training", self.conf.image_downsample)
        i = self.model.modalities.index(modality)
        return self.get_data_generator(train_images=[d.get_images_modi(i)], train_labels=None)

    # ------------------------------------------------------------------ staging helpers
    def _stage(self, gen):
        """next(gen) -> device tensors (async copies from the pinned buffers)"""
        item = next(gen)
        items = list(item) if isinstance(item, tuple) else [item]
        n = min(t.shape[0] for t in items)
        out = []
        for t in items:
            t = t[:n]
            self.h2d_bytes += t.numel() * 4
            out.append(t.cuda(non_blocking=True))
        gen.mark_copied()
        theta = getattr(gen, "last_theta", None)
        if theta is not None:
            # augmentation (base_executor.py:103-110) on the device: the same angles for every array of the group
            th = torch.from_numpy(theta[:n]).cuda(non_blocking=True)
            self.h2d_bytes += th.numel() * 4
            out = [ops.rotate_bilinear(t.contiguous(), th) for t in out]
            for i in getattr(gen, "residual_items", ()):
                ops.mask_residual_(out[i])          # background channel of the rotated masks (dafnet_executor.py:493-494)
        return out

    def _sample_z(self, B):
        norm = NormalDistribution()
        z = norm.sample((B, self.conf.num_z)).astype(np.float32)
        self.h2d_bytes += z.size * 4
        return torch.from_numpy(z).cuda(non_blocking=True)

    def _sample_idx(self, n, k):
        idx = data_utils.sample_indices(n, k).astype(np.int32)
        self.h2d_bytes += idx.size * 4
        return torch.from_numpy(idx).cuda(non_blocking=True)

    # ------------------------------------------------------------------ training loop
    def get_loss_names(self):
        """dafnet_executor.py:200-205 (the columns of training.csv, in the reference's order)"""
        return ["adv_M", "adv_X1", "adv_X2", "rec_X", "dis_M", "dis_X1", "dis_X2", "val_loss", "val_loss_mod1", "val_loss_mod2",
                "val_loss_mod2_mod1def", "val_loss_mod1_mod2def", "val_loss_mod2_fused", "val_loss_mod1_fused",
                "val_weight_0", "val_weight_1", "val_weight_2", "supervised_Mask", "KL", "rec_Z"]

    def train(self):
        log.info("Training Model")
        self.init_train_data()
        os.makedirs(self.conf.folder, exist_ok=True)
        csv_path = os.path.join(self.conf.folder, "training.csv")
        names = self.get_loss_names()
        best, wait = None, 0
        with open(csv_path, "w") as f:
            f.write("epoch," + ",".join(names) + "\n")
        for self.epoch in range(self.conf.epochs):
            log.info("Epoch %d/%d" % (self.epoch, self.conf.epochs))
            epoch_loss = {n: [] for n in names}
            for self.batch in range(self.batches):
                self.train_batch(epoch_loss)
            self.flush_losses(epoch_loss)
            self.set_swa_model_weights()
            for swa_m in self.get_swa_models():
                swa_m.on_epoch_end(self.epoch)
            self.validate(epoch_loss)
            row = [np.mean(epoch_loss[n]) if len(epoch_loss[n]) else 0.0 for n in names]
            with open(csv_path, "a") as f:
                f.write(str(self.epoch) + "," + ",".join("%.6g" % v for v in row) + "\n")
            self.save_models()
            # EarlyStopping(min_delta=0.01, patience=60) on val_loss_mod2_fused (dafnet_executor.py:222,263-284)
            cur = row[names.index("val_loss_mod2_fused")]
            if best is None or cur < best - 0.01:
                best, wait = cur, 0
            else:
                wait += 1
                if wait >= 60:
                    log.info("Finished training from early stopping criterion")
                    # final model parameters = the stochastic weight average (dafnet_executor.py:268-284)
                    for swa_m in self.get_swa_models():
                        swa_m.on_train_end()
                    self.save_models()
                    break

    def validate(self, epoch_loss):
        """dafnet_executor.py:303-367: 1 - Dice(binarised) on the validation split for both modalities -- from their own
        anatomy, from the other modality's anatomy deformed onto them and from the fused anatomy -- through the SWA clones;
        with automated pairing also the Balancer's mean weight per candidate (live models, as the reference)."""
        valid = self.loader.load_all_modalities_concatenated(self.conf.split, "validation", self.conf.image_downsample)
        if getattr(self.conf, "randomise", False):
            valid.randomise_pairs(length=self.conf.n_pairs - 1)
        valid.crop(self.conf.input_shape[:2])
        nm = self.loader.num_masks
        x0, x1 = valid.get_images_modi(0), valid.get_images_modi(1)
        real0, real1 = valid.get_masks_modi(0)[..., :nm], valid.get_masks_modi(1)[..., :nm]
        # up to SWA_EPOCH the clones' weights ARE the live weights, so the clones are only built once averaging has started
        averaging = (self.USE_SWA and getattr(self, "epoch", 0) > self.SWA_EPOCH
                     and self.swa_Segmentor.swa_weights is not None)
        pick = (lambda name, live: getattr(self, name).get_clone_model()) if averaging else (lambda name, live: live)
        enc0 = pick("swa_Enc_Anatomy1", self.model.Encoders_Anatomy[0])
        enc1 = pick("swa_Enc_Anatomy2", self.model.Encoders_Anatomy[1])
        seg = pick("swa_Segmentor", self.model.Segmentor)
        fuser = pick("swa_Anatomy_Fuser", self.model.Anatomy_Fuser)
        s1 = enc0.predict(x0)
        s2 = enc1.predict(x1)
        s1_deformed, s2_fused = fuser.predict([s1, s2])
        s2_deformed, s1_fused = fuser.predict([s2, s1])
        loss = lambda real, s: 1 - costs.dice(real, seg.predict(s), binarise=True)
        dice_m1s1, dice_m1s2def, dice_m1fused = loss(real0, s1), loss(real0, s2_deformed), loss(real0, s1_fused)
        dice_m2s2, dice_m2s1def, dice_m2fused = loss(real1, s2), loss(real1, s1_deformed), loss(real1, s2_fused)
        epoch_loss["val_loss_mod2"].append(dice_m2s2)
        epoch_loss["val_loss_mod2_mod1def"].append(dice_m2s1def)
        epoch_loss["val_loss_mod2_fused"].append(dice_m2fused)
        epoch_loss["val_loss_mod1_mod2def"].append(dice_m1s2def)
        epoch_loss["val_loss_mod1_fused"].append(dice_m1fused)
        epoch_loss["val_loss_mod1"].append(dice_m1s1)
        epoch_loss["val_loss"].append(np.mean([dice_m1s1, dice_m2s2, dice_m2s1def, dice_m2fused]))
        if getattr(self.conf, "automatedpairing", False):
            valid.expand_pairs(self.conf.n_pairs - 1, 0, neighborhood=self.conf.n_pairs)
            x0p = valid.get_images_modi(0)
            s1_list = [self.model.Encoders_Anatomy[0].predict(np.ascontiguousarray(x0p[..., i:i + 1]))
                       for i in range(x0p.shape[-1])]
            s2_live = self.model.Encoders_Anatomy[1].predict(valid.get_images_modi(1))
            weights = self.model.Balancer.predict([s2_live] + s1_list)
            for j in range(weights.shape[-1]):
                epoch_loss.setdefault("val_weight_%d" % j, []).append(float(np.mean(weights[..., j])))

    def train_batch(self, epoch_loss):
        """dafnet_executor.py:369-387"""
        if self._graph is not None:
            # the snapshots of the previous replay must be read before they are overwritten
            self.flush_losses(epoch_loss)
            step = self._next if self._next is not None else self.stage_step_inputs()
            if self._next_ready is not None:
                # the batch was staged on the copy stream: the replay's input copies wait for it, and the allocator must
                # not hand the staged tensors' memory to the copy stream again while this stream still reads them
                main = torch.cuda.current_stream()
                main.wait_event(self._next_ready)
                for t in self._flat(step):
                    t.record_stream(main)
                self._next_ready = None
            self.train_batch_graph(step)
            # prefetch: gather the next batches into pinned memory and run their H2D copies (and the rotation kernels) on a
            # COPY stream while the GPU is busy with the replay that was just launched -- enqueued on the compute stream
            # they would only start after the replay (154 MB = ~3 ms per step at 224^2 x 32 pairs)
            self._next = self._stage_ahead()
            return
        if self.conf.l_mix > 0:
            self.train_supervised_expert_pairing(epoch_loss)
            self.train_batch_mask_discriminator(epoch_loss)
            self.train_batch_image_discriminator(epoch_loss)
        if self.conf.l_mix < 1:
            self.train_unsupervised_expert_pairing(epoch_loss)
            self.train_batch_mask_discriminator(epoch_loss)
            self.train_batch_image_discriminator(epoch_loss)

    # -- staging (host -> device) is separated from the device work so that a pre-staged step can be replayed
    def _stage_generator(self, supervised):
        batch = self._stage(self.gen_labelled if supervised else self.gen_unlabelled)
        B = batch[0].shape[0]
        return batch + [self._sample_z(B) for _ in range(4)]       # z1, z2 (sampled inputs), eps1, eps2

    def _stage_mask_d(self):
        (m1,) = self._stage(self.discriminator_masks)
        (m2,) = self._stage(self.discriminator_masks)
        (x1,) = self._stage(self.discriminator_image[0])
        (x2,) = self._stage(self.discriminator_image[1])
        B = min(t.shape[0] for t in (x1, x2, m1, m2))
        nm = self.conf.num_masks                 # dafnet_executor.py:516-517: [..., 0:conf.num_masks]
        return [x1[:B], x2[:B], m1[:B, ..., 0:nm].contiguous(), m2[:B, ..., 0:nm].contiguous(), self._sample_idx(2 * B, B),
                self._sample_idx(2 * B, B)]

    def _stage_image_d(self):
        (x1,) = self._stage(self.discriminator_image[0])
        (x2,) = self._stage(self.discriminator_image[1])
        B = min(x1.shape[0], x2.shape[0])
        return [x1[:B].contiguous(), x2[:B].contiguous(), self._sample_z(B), self._sample_z(B),
                self._sample_idx(3 * B, B), self._sample_idx(3 * B, B)]

    def stage_step_inputs(self):
        """everything one train_batch reads from the host, as device tensors"""
        step = []
        if self.conf.l_mix > 0:
            step.append(("sup", self._stage_generator(True), self._stage_mask_d(), self._stage_image_d()))
        if self.conf.l_mix < 1:
            step.append(("unsup", self._stage_generator(False), self._stage_mask_d(), self._stage_image_d()))
        return step

    def _stage_ahead(self):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        with torch.cuda.stream(self._copy_stream):
            step = self.stage_step_inputs()
            self._next_ready = torch.cuda.Event()
            self._next_ready.record()
        return step

    def train_batch_on(self, step):
        """train_batch on pre-staged (HBM-resident) inputs"""
        if self._graph is not None:
            return self.train_batch_graph(step)
        self._run_step(step)

    def _run_step(self, step):
        """the device work of one train_batch on staged inputs (entries whose part is None are skipped)"""
        for kind, g, dm, di in step:
            if g is not None:
                self._run_generator(kind == "sup", g)
            if dm is not None:
                self._run_mask_d(dm)
            if di is not None and len(di):
                self._run_image_d(di)

    @staticmethod
    def _flat(step):
        """all staged device tensors of a step, in a fixed order"""
        return [t for _, g, dm, di in step for part in (g, dm, di) if part is not None for t in part]

    # ------------------------------------------------------------------ CUDA-graph replay of the whole step
    def enable_cuda_graph(self, warmup=2):
        """Capture one whole train_batch (every kernel of the generator update, the two mask-discriminator updates
        and the two image-discriminator updates, ~3000 launches) into ONE CUDA graph that reads its inputs from
        static device buffers.  A step is then: H2D of the new batch -> copy into the static buffers -> one graph
        launch.  Everything the step needs between replays lives on the device (Adam step count / lr_t, BatchNorm
        moving statistics, packed bf16 weight copies, loss slots)."""
        assert self._graph is None
        self._static = self.stage_step_inputs()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):          # steady state: allocator warm, weights packed, attributes set
                self._pending = []
                self._run_step(self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._pending = []
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._run_step(self._static)
        self._graph_pending = list(self._pending)     # loss snapshots live in the graph's memory pool
        self._pending = []
        self._graph = graph
        return graph

    def train_batch_graph(self, step):
        if step is not self._static:
            src_l, dst_l = self._flat(step), self._flat(self._static)
            same = len(src_l) == len(dst_l) and all(tuple(a.shape) == tuple(b.shape) for a, b in zip(src_l, dst_l))
            if not same:          # ragged last batch of an epoch: run this one step eagerly
                graph, self._graph = self._graph, None
                try:
                    return self.train_batch_on(step)
                finally:
                    self._graph = graph
            for src, dst in zip(src_l, dst_l):
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        self._pending = list(self._graph_pending)

    def train_supervised_expert_pairing(self, epoch_loss):
        """dafnet_executor.py:389-411"""
        self._run_generator(True, self._stage_generator(True))

    def train_unsupervised_expert_pairing(self, epoch_loss):
        """dafnet_executor.py:413-434"""
        self._run_generator(False, self._stage_generator(False))

    def train_batch_mask_discriminator(self, epoch_loss):
        """dafnet_executor.py:511-545"""
        self._run_mask_d(self._stage_mask_d())

    def train_batch_image_discriminator(self, epoch_loss):
        """dafnet_executor.py:547-583"""
        self._run_image_d(self._stage_image_d())

    def _run_generator(self, supervised, t):
        if getattr(self.conf, "automatedpairing", False):
            return self._run_generator_paired(supervised, t)
        if supervised:
            x1, x2, m1, m2, z1, z2, eps1, eps2 = t
            tr = self.model.supervised_trainer
            tr.train_on_device(x1, x2, z1, z2, eps1, eps2, m1, m2)
        else:
            x1, x2, m1, z1, z2, eps1, eps2 = t
            tr = self.model.unsupervised_trainer
            tr.train_on_device(x1, x2, z1, z2, eps1, eps2, m1)
        self._pending.append((tr, tr.book.snapshot(), "gen"))

    def _run_generator_paired(self, supervised, t):
        """train_{un,}supervised_automated_pairing + prepare_data_to_train (dafnet_executor.py:436-499): the staged
        images carry n_pairs candidates on the channel axis; candidate 0 is the expert pair"""
        P = int(self.conf.n_pairs)
        x1p, x2p = t[0], t[1]
        split = lambda x: [ops.slice_channels(x, i, 1) for i in range(P)]
        x1_lst, x2_lst = split(x1p), split(x2p)
        if supervised:
            m1, m2, z1, z2, eps1, eps2 = t[2:]
            tr = self.model.supervised_trainer
            tr.train_on_device(*(x1_lst + x2_lst + [z1, z2, eps1, eps2, m1, m2]))
        else:
            m1, z1, z2, eps1, eps2 = t[2:]
            tr = self.model.unsupervised_trainer
            tr.train_on_device(*(x1_lst + x2_lst + [z1, z2, eps1, eps2, m1]))
        self._pending.append((tr, tr.book.snapshot(), "gen"))

    def train_supervised_automated_pairing(self, epoch_loss):
        self._run_generator(True, self._stage_generator(True))

    def train_unsupervised_automated_pairing(self, epoch_loss):
        self._run_generator(False, self._stage_generator(False))

    def _run_mask_d(self, t):
        x1, x2, m1, m2, idx1, idx2 = t
        nm = self.conf.num_masks
        B = x1.shape[0]
        M = self.model
        fake_s1 = M.Encoders_Anatomy[0].predict_device(x1)
        fake_s2 = M.Encoders_Anatomy[1].predict_device(x2)
        for real, idx, s_own, s_a, s_b in ((m1, idx1, fake_s1, fake_s2, fake_s1), (m2, idx2, fake_s2, fake_s1, fake_s2)):
            fake_m = M.Segmentor.predict_device(s_own)
            s_def = M.Anatomy_Fuser.predict_deform_device(s_a, s_b)
            fake_m_def = M.Segmentor.predict_device(s_def)
            cat = torch.empty((2 * B,) + tuple(fake_m.shape[1:3]) + (nm,), dtype=torch.float32, device="cuda")
            ops.copy_channels(fake_m, 0, cat[:B], 0, nm)
            ops.copy_channels(fake_m_def, 0, cat[B:], 0, nm)
            fake = ops.gather_rows(cat, idx)
            M.D_Mask_trainer.train_on_device(real, fake)
            self._pending.append((M.D_Mask_trainer, M.D_Mask_trainer.book.snapshot(), "dis_M"))

    def _run_image_d(self, t):
        x1, x2, eps1, eps2, idx1, idx2 = t
        M = self.model
        s1 = M.Encoders_Anatomy[0].predict_device(x1)
        s2 = M.Encoders_Anatomy[1].predict_device(x2)
        s1_def = M.Anatomy_Fuser.predict_deform_device(s1, s2)
        s2_def = M.Anatomy_Fuser.predict_deform_device(s2, s1)
        z1 = self._predict_z(s1, x1, eps1)
        z2 = self._predict_z(s2, x2, eps2)
        # six Decoder.predict calls (dafnet_executor.py:560-570) as one batched call: the decoder has no batch statistics
        B = x1.shape[0]
        ys = M.Decoder.predict_device(_cat_rows([s1, s2_def, s1_def, s2, s1_def, s2_def]), _cat_rows([z1, z1, z1, z2, z2, z2]))
        y1 = ops.gather_rows(ys[:3 * B], idx1)
        y2 = ops.gather_rows(ys[3 * B:], idx2)
        M.D_Image1_trainer.train_on_device(x1, y1)
        self._pending.append((M.D_Image1_trainer, M.D_Image1_trainer.book.snapshot(), "dis_X1"))
        M.D_Image2_trainer.train_on_device(x2, y2)
        self._pending.append((M.D_Image2_trainer, M.D_Image2_trainer.book.snapshot(), "dis_X2"))

    def _predict_z(self, s, x, eps):
        """Enc_Modality.predict -> z (inference phase)"""
        mu, lv = self.model.Enc_Modality.predict_device(s, x)
        z, _ = ops.vae_fwd(mu, lv, eps, 0.0, None)
        return z

    # ------------------------------------------------------------------ loss bookkeeping
    def flush_losses(self, epoch_loss):
        """one D2H read per trainer call; kept out of the step so the device never waits for the host"""
        for tr, snap, kind in self._pending:
            h = tr.book.history(snap)
            self.d2h_bytes += snap.numel() * 4
            if kind == "gen":
                epoch_loss["supervised_Mask"].append(h["Segmentor_loss"][0])
                epoch_loss["adv_M"].append(h["D_Mask_loss"][0])
                epoch_loss["rec_X"].append(h["Decoder_loss"][0])
                epoch_loss["adv_X1"].append(h["D_Image1_loss"][0])
                epoch_loss["adv_X2"].append(h["D_Image2_loss"][0])
                epoch_loss["KL"].append(h["Enc_Modality_loss"][0])
                epoch_loss["rec_Z"].append(h["ZReconstruct_loss"][0])
                epoch_loss.setdefault("loss", []).append(h["loss"][0])       # sum of the weighted terms (not a column of the
                #                                                              reference's DAFNet csv; MMSDNet lists it)
            else:
                epoch_loss[kind].append(h["loss"][0])
        self._pending = []


def _cat_rows(ts):
    """concatenate along the batch axis with our own copy kernel"""
    n = sum(t.shape[0] for t in ts)
    out = torch.empty((n,) + tuple(ts[0].shape[1:]), dtype=ts[0].dtype, device="cuda")
    off = 0
    for t in ts:
        ops.copy_(out[off:off + t.shape[0]], t)
        off += t.shape[0]
    return out

"""Base executor (reference: model_executors/base_executor.py:14-119).

Data generators: the reference wraps every array in a Keras ImageDataGenerator (rotation_range 20,
base_executor.py:103-110).  Augmentation is out of scope for the hot path (synthetic inputs), so a
generator here is a seeded, shuffling mini-batch iterator that stages each batch in PINNED host
memory -- the host->device copy of the step is an asynchronous cudaMemcpy from that buffer.
"""
import logging

import numpy as np
import torch

from ..loaders import loader_factory

log = logging.getLogger("executor")


class BatchFlow(object):
    """equivalent of ImageDataGenerator().flow(array, batch_size, seed) without augmentation"""

    def __init__(self, array, batch_size, seed, rotation_range=0.0):
        self.array = np.ascontiguousarray(array, dtype=np.float32)
        self.batch_size = batch_size
        self.rng = np.random.RandomState(seed)
        # ImageDataGenerator(rotation_range): one angle per sample, theta = deg2rad(uniform(-range, range)); flows built
        # with the same seed draw the same angles, which keeps images and masks aligned (base_executor.py:37-78).
        # The rotation itself runs on the device after the H2D copy (ops.rotate_bilinear).
        self.rotation_range = float(rotation_range)
        self.last_theta = None
        self.n = self.array.shape[0]
        self.order = None
        self.pos = 0
        self.pinned = [None, None]      # double-buffered pinned staging
        self.events = [None, None]      # H2D-complete events recorded by the consumer
        self.turn = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self.order is None or self.pos >= self.n:
            self.order = self.rng.permutation(self.n)
            self.pos = 0
        idx = self.order[self.pos:self.pos + self.batch_size]
        self.pos += self.batch_size
        if self.rotation_range:
            self.last_theta = np.deg2rad(self.rng.uniform(-self.rotation_range, self.rotation_range,
                                                          size=len(idx))).astype(np.float32)
        shape = (len(idx),) + self.array.shape[1:]
        if torch.cuda.is_available():
            t = self.turn
            self.turn ^= 1
            if self.pinned[t] is None:
                self.pinned[t] = torch.empty((self.batch_size,) + self.array.shape[1:], dtype=torch.float32).pin_memory()
            if self.events[t] is not None:
                self.events[t].synchronize()     # the copy that last read this buffer has finished
            out = self.pinned[t][:len(idx)]
            np.take(self.array, idx, axis=0, out=out.numpy())
            self._last = t
            return out
        return torch.from_numpy(self.array[idx].reshape(shape))

    def mark_copied(self):
        """called by the consumer right after it enqueued the H2D copy of the last batch"""
        if torch.cuda.is_available():
            ev = torch.cuda.Event()
            ev.record()
            self.events[self._last] = ev


class FlowGroup(object):
    """zip of aligned BatchFlows (same seed -> same order), yields a tensor or a tuple of tensors"""

    def __init__(self, flows, residual_items=()):
        self.flows = flows
        # positions of the mask arrays whose last channel is the add_residual background: the reference builds it AFTER
        # the augmentation (dafnet_executor.py:493-494), so the stager rebuilds it once the batch has been rotated
        self.residual_items = tuple(residual_items)

    def __iter__(self):
        return self

    def __next__(self):
        items = [next(f) for f in self.flows]
        return items[0] if len(items) == 1 else tuple(items)

    @property
    def last_theta(self):
        """rotation angles of the batch just produced (identical in every flow of the group), or None"""
        return self.flows[0].last_theta

    def mark_copied(self):
        for f in self.flows:
            f.mark_copied()


class Executor(object):
    def __init__(self, conf, model):
        self.conf = conf
        self.model = model
        self.loader = loader_factory.init_loader(self.conf.dataset_name)
        if hasattr(self.loader, "input_shape") and hasattr(conf, "input_shape"):
            self.loader.input_shape = tuple(conf.input_shape)     # synthetic data follows the configured size
        self.batch = 0
        self.epoch = 0

    def init_train_data(self):
        pass

    def get_loss_names(self):
        pass

    def train(self):
        pass

    def get_data_generator(self, train_images=None, train_labels=None, labels_have_residual=False):
        """base_executor.py:37-78: zip of one flow per array, all seeded with conf.seed so that the image and
        label streams stay aligned"""
        gens, residual = [], []
        for is_label, arrs in ((False, train_images), (True, train_labels)):
            if arrs is None:
                continue
            if type(arrs) != list:
                arrs = [arrs]
            for a in arrs:
                if labels_have_residual and is_label:
                    residual.append(len(gens))
                gens.append(BatchFlow(a, self.conf.batch_size, self.conf.seed,
                                      self.get_datagen_params()["rotation_range"] if getattr(self.conf, "augment", True) else 0.0))
        if len(gens) == 0:
            raise Exception("No data to iterate.")
        return FlowGroup(gens, residual)

    def validate(self, epoch_loss):
        pass

    def add_residual(self, data):
        """base_executor.py:83-87: background channel = 1 where no mask channel is set"""
        residual = np.ones(data.shape[:-1] + (1,), dtype=data.dtype)
        for i in range(data.shape[-1]):
            residual[data[..., i:i + 1] == 1] = 0
        return np.concatenate([data, residual], axis=-1)

    def test(self):
        from ..model_tester import ModelTester
        log.info("Evaluating model on test data")
        tester = ModelTester(self.model, self.conf)
        tester.run()

    def get_datagen_params(self):
        return dict(horizontal_flip=False, vertical_flip=False, rotation_range=20.,
                    width_shift_range=0, height_shift_range=0, zoom_range=0)

    def align_batches(self, array_list):
        """base_executor.py:112-119"""
        mn = np.min([x.shape[0] for x in array_list])
        return [x[0:mn] for x in array_list]

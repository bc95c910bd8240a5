"""Trainer plumbing shared by MMSDNet and DAFNet: the compiled-``Model`` behaviour of Keras that the
executors rely on (``fit(inputs, targets)`` = exactly one Adam step per <= 32 samples, returning a
``History``), implemented as: stage inputs -> zero the flat gradient bucket -> forward on the tape
-> fused loss kernels seed the gradients -> tape backward -> (NCCL all-reduce of the bucket when
data-parallel) -> one fused Adam launch.
"""
import numpy as np
import torch

from .. import engine as E
from .. import ops


class History(object):
    def __init__(self, history):
        self.history = history


def to_dev(a):
    if a is None:
        return None
    if torch.is_tensor(a):
        return a if a.is_cuda else a.cuda(non_blocking=True)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda(non_blocking=True)


class LossBook(object):
    """one fp32 slot per model output (weighted value), filled by the loss kernels on the device"""

    def __init__(self, names_weights):
        self.names = [n for n, _ in names_weights]
        self.weights = [float(w) for _, w in names_weights]
        self.buf = None

    def reset(self):
        if self.buf is None:
            self.buf = ops.zeros(max(len(self.names), 1))
        else:
            ops.zero_(self.buf)

    def slot(self, i):
        return self.buf[i:i + 1]

    def snapshot(self):
        """device copy of the slots (our own copy kernel), so the read can be deferred"""
        return ops.cast(self.buf, torch.float32)

    def history(self, vals=None):
        """Keras History semantics: 'loss' = weighted total; '<name>_loss' = unweighted value of the LAST
        output carrying that name (dafnet_executor.py:502-509 reads exactly those)."""
        src = self.buf if vals is None else vals
        vals = src.detach().cpu().numpy().astype(np.float64)
        h = {"loss": [float(vals[:len(self.names)].sum())]}
        for n, w, v in zip(self.names, self.weights, vals):
            h[n + "_loss"] = [float(v / w) if w != 0 else 0.0]
        return h


class Trainer(object):
    """Base: owns an Adam state over ``params`` and runs ``graph(ctx, book, *device_inputs)``."""
    dist = None      # set by parallel.DataParallel: object with allreduce_(list_of_buckets) and world_size

    def __init__(self, name, params, lr, names_weights, frozen_models=()):
        self.name = name
        self.params = params
        self.opt = E.Adam(params, lr)
        self.book = LossBook(names_weights)
        # models that take part in the graph with frozen weights (make_trainable(D, False),
        # models/dafnet.py:119-121): gradients flow through them but not into them
        self.frozen_models = list(frozen_models)

    def graph(self, ctx, book, *inputs):
        raise NotImplementedError

    def extra_grads(self, book):
        """hook: regulariser gradients that do not flow through the tape (Spectral)"""

    def train_on_device(self, *inputs):
        self.forward_backward(*inputs)
        self.apply_gradients()

    AR_CHUNK = 8 << 20      # floats per all-reduce piece (32 MB): large enough for NVLink bandwidth, small enough to overlap

    def forward_backward(self, *inputs):
        """loss + gradients into the flat bucket (no weight update).  Data-parallel: a piece of the gradient arena is
        all-reduced (asynchronously, on NCCL's stream) as soon as the last node of the backward pass that writes into it
        has run, so the collective overlaps the rest of the backward pass; apply_gradients() waits for the pieces."""
        self.book.reset()
        self.opt.zero_grad()
        tape = E.Tape()
        ctx = E.Ctx(tape, training=True)
        for m in self.frozen_models:
            m.trainable = False
        self._ar_works, self._ar_done = [], set()
        try:
            self.graph(ctx, self.book, *inputs)
            after = None
            d = Trainer.dist
            overlap = (d is not None and d.world_size > 1 and getattr(d, "overlap", False)
                       and type(self).extra_grads is Trainer.extra_grads)     # regulariser gradients are added after the tape
            if overlap:
                pieces = self._ar_pieces()
                first = tape.last_writers(pieces)
                by_node = {}
                for k, i in enumerate(first):
                    if i is not None:
                        by_node.setdefault(i, []).append(k)

                def after(i):
                    for k in by_node.get(i, ()):
                        arena, a, b = pieces[k]
                        self._ar_works.append(d.allreduce_async(arena.gflat[a:b]))
                        self._ar_done.add((id(arena), a, b))
            tape.backward(after)
        finally:
            for m in self.frozen_models:
                m.trainable = True
        self.extra_grads(self.book)

    def _ar_pieces(self):
        out = []
        for arena, a, b in self.opt.ranges:
            for lo in range(a, b, self.AR_CHUNK):
                out.append((arena, lo, min(b, lo + self.AR_CHUNK)))
        return out

    def apply_gradients(self):
        """(all-reduce when data-parallel) + one fused Adam launch per bucket"""
        scale = 1.0
        d = Trainer.dist
        if d is not None and d.world_size > 1:
            done = getattr(self, "_ar_done", set())
            if done:
                rest = [arena.gflat[a:b] for arena, a, b in self._ar_pieces() if (id(arena), a, b) not in done]
                d.allreduce_(rest)
                d.wait(self._ar_works)
                self._ar_works, self._ar_done = [], set()
            else:
                d.allreduce_(self.opt.grad_buckets())
            scale = 1.0 / d.world_size
        self.opt.step(grad_scale=scale)

    def fit(self, inputs, targets, epochs=1, verbose=0, batch_size=32):
        """numpy in; Keras splits into batches of 32 = one optimizer step each."""
        inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        targets = list(targets) if isinstance(targets, (list, tuple)) else [targets]
        n = inputs[0].shape[0]
        hist, sizes = [], []
        for s in range(0, n, batch_size):
            di = [to_dev(np.asarray(a)[s:s + batch_size]) for a in inputs]
            dt = [to_dev(np.asarray(a)[s:s + batch_size]) if isinstance(a, np.ndarray) and a.ndim > 0 and a.shape[0] == n
                  else a for a in targets]
            self.train_on_device(*self.pack(di, dt))
            hist.append(self.book.history())
            sizes.append(di[0].shape[0])
        out = {}
        for k in hist[0]:
            out[k] = [float(np.average([h[k][0] for h in hist], weights=sizes))]
        return History(out)

    def pack(self, dev_inputs, dev_targets):
        """map Keras-style (inputs, targets) lists onto the positional arguments of ``graph``"""
        return list(dev_inputs) + list(dev_targets)


class DiscriminatorTrainer(Trainer):
    """models/mmsdnet.py:62-77, models/dafnet.py:75-115: Model([real, fake], [D(real), D(fake)]),
    loss 'mse' against [ones, zeros], plus the Spectral regularisation losses of the D layers."""

    def __init__(self, name, D, lr):
        super().__init__(name, D.params(), lr, [(D.name + "_real", 1.0), (D.name + "_fake", 1.0), (D.name + "_reg", 1.0)])
        self.D = D

    def graph(self, ctx, book, real, fake):
        pr = self.D(ctx, E.Var(real))
        pf = self.D(ctx, E.Var(fake))
        E.loss_l1l2(ctx, pr, None, 1, 1.0, book.slot(0), cval=1.0)
        E.loss_l1l2(ctx, pf, None, 1, 1.0, book.slot(1), cval=0.0)

    def extra_grads(self, book):
        for conv, reg in getattr(self.D, "regularizers", []):
            reg(conv.kernel, book.slot(2), with_grad=True)

    def pack(self, dev_inputs, dev_targets):
        return dev_inputs[:2]

    def fit(self, inputs, targets, epochs=1, verbose=0, batch_size=32):
        h = super().fit(inputs, targets, epochs, verbose, batch_size)
        # Keras names: the two outputs share the model name; 'loss' is the total incl. regularisers
        h.history[self.D.name + "_loss"] = h.history[self.D.name + "_fake_loss"]
        return h

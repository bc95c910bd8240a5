"""The slice of the Keras object protocol that the reference's executors use:
``Model.predict / get_weights / set_weights / save_weights / load_weights / summary /
output_shape / get_output_shape_at / name / trainable``, plus the build scope that lets the
reference-style ``build(conf)`` functions place their parameters into a shared flat arena.
"""
import os
import threading

import numpy as np
import torch

from . import engine as E


class EasyDict(dict):
    """minimal easydict.EasyDict (attribute access, recursive); the package is not installed here."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


# ------------------------------------------------------------------------------------------------
class BuildScope:
    """``with BuildScope(arena, state, rng):`` -- builders called inside add their parameters to the
    given arenas (owner finalises).  Outside any scope a builder gets private arenas."""
    _tls = threading.local()

    def __init__(self, arena=None, state=None, rng=None, seed=0):
        self.arena = arena if arena is not None else E.Arena(True)
        self.state = state if state is not None else E.Arena(False)
        self.rng = rng if rng is not None else np.random.RandomState(seed)
        self.private = False

    def __enter__(self):
        stack = getattr(self._tls, "stack", None)
        if stack is None:
            stack = self._tls.stack = []
        stack.append(self)
        return self

    def __exit__(self, *a):
        self._tls.stack.pop()

    @classmethod
    def current(cls):
        stack = getattr(cls._tls, "stack", None)
        if stack:
            return stack[-1]
        s = BuildScope()
        s.private = True
        return s


class InputSpec(object):
    """one entry of ``Model.inputs`` (keras Input tensor stand-in): the owning model, its position and shape"""

    def __init__(self, model, index, shape):
        self.model, self.index, self.shape = model, index, (None,) + tuple(shape)

    def __repr__(self):
        return "<Input %d of %s %s>" % (self.index, self.model.name, self.shape)


class LayerOutput(object):
    """``model.get_layer(name).output``: the value of an inner layer, usable as the output of a sub-model
    ``Model(model.inputs, model.get_layer('z_mean').output)`` (models/dafnet.py:126, models/mmsdnet.py:51,87)"""

    def __init__(self, model, layer):
        self.model, self.layer = model, layer


class LayerHandle(object):
    """``model.get_layer(name)``: name, weights and ``.output`` of one layer of a component"""

    def __init__(self, model, layer):
        self.model, self.layer, self.name = model, layer, layer.name

    def _weights(self):
        ws = list(self.layer.params())
        if isinstance(self.layer, E.BatchNorm):
            ws += [self.layer.moving_mean, self.layer.moving_var]
        return ws

    def get_weights(self):
        self.model._ensure_device()
        torch.cuda.synchronize() if torch.cuda.is_available() else None
        return [p.numpy() for p in self._weights()]

    def set_weights(self, weights):
        self.model._ensure_device()
        ws = self._weights()
        assert len(ws) == len(weights), "%s: expected %d arrays, got %d" % (self.name, len(ws), len(weights))
        for p, w in zip(ws, weights):
            w = np.asarray(w, np.float32)
            assert tuple(w.shape) == p.shape, (p.name, w.shape, p.shape)
            p.data.copy_(torch.from_numpy(np.ascontiguousarray(w)))
        self.model._scope.arena.version += 1
        self.model._scope.state.version += 1

    @property
    def output(self):
        if self.name not in self.model._taps:
            raise ValueError("layer %r of %s has no registered output tap (taps: %s)"
                             % (self.name, self.model.name, sorted(self.model._taps)))
        return LayerOutput(self.model, self.layer)


class Model:
    """A named component: an ordered list of layers and a forward function over ``Var``s.

    Two constructors, as in Keras: the builders' ``Model(name, layers, forward, input_shapes, output_shapes, scope)`` and
    the functional ``Model(model.inputs, model.get_layer(name).output[, name=...])`` for a sub-graph that shares the
    parent's layers (and therefore its weights)."""

    def __init__(self, *args, **kw):
        first = args[0] if args else kw.get("inputs")
        if isinstance(first, (list, tuple)) and len(first) > 0 and isinstance(first[0], InputSpec):
            inputs = first
            out = args[1] if len(args) > 1 else kw.get("outputs")
            if not isinstance(out, LayerOutput):
                raise TypeError("Model(inputs, outputs): outputs must be model.get_layer(name).output")
            parent = out.model
            if list(inputs) != list(parent.inputs):
                raise ValueError("Model(inputs, outputs): inputs must be the parent model's .inputs")
            forward, layers, oshape = parent._taps[out.layer.name]
            name = kw.get("name") or (parent.name + "_" + out.layer.name)
            input_shapes, output_shapes, scope = parent.input_shapes, [oshape], parent._scope
        else:
            name, layers, forward, input_shapes, output_shapes, scope = args
        self.name = name
        self.layers = layers                  # creation order == Keras weight order
        self._forward = forward
        self.input_shapes = input_shapes      # without the batch axis
        self.output_shapes = output_shapes
        self._scope = scope
        self._trainable = True
        self._taps = {}                       # layer name -> (forward up to that layer, its layers, output shape)
        self._inputs = None

    # -- device placement ------------------------------------------------------------------
    def _ensure_device(self):
        sc = self._scope
        if sc.arena.flat is None:
            sc.arena.to_device()
        if sc.state.flat is None:
            sc.state.to_device()

    # -- graph use ---------------------------------------------------------------------------
    def __call__(self, ctx, *inputs):
        self._ensure_device()
        return self._forward(ctx, *inputs)

    # -- keras protocol ------------------------------------------------------------------------
    @property
    def inputs(self):
        if self._inputs is None:
            self._inputs = [InputSpec(self, i, s) for i, s in enumerate(self.input_shapes)]
        return self._inputs

    @property
    def input_shape(self):
        shp = [(None,) + tuple(s) for s in self.input_shapes]
        return shp[0] if len(shp) == 1 else shp

    def register_tap(self, layer_name, forward, layers, output_shape):
        """make ``get_layer(layer_name).output`` usable as the output of a sub-model"""
        self._taps[layer_name] = (forward, layers, tuple(output_shape))

    def get_layer(self, name=None, index=None):
        """keras ``Model.get_layer``: by name, or by position in the layer list"""
        if index is not None:
            return LayerHandle(self, self.layers[index])
        for l in self.layers:
            if getattr(l, "name", None) == name:
                return LayerHandle(self, l)
        raise ValueError("No such layer: %s" % name)

    @property
    def output_shape(self):
        shp = [(None,) + tuple(s) for s in self.output_shapes]
        return shp[0] if len(shp) == 1 else shp

    def get_output_shape_at(self, idx):
        return self.output_shape

    @property
    def trainable(self):
        return self._trainable

    @trainable.setter
    def trainable(self, val):
        self._trainable = bool(val)
        for p in self.params():
            p.requires_grad = bool(val)

    def params(self):
        out, seen = [], set()
        for l in self.layers:
            for p in l.params():
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out

    def weight_list(self):
        """Keras ordering: per layer, trainable weights then non-trainable (BN moving stats)."""
        out, seen = [], set()
        for l in self.layers:
            ws = list(l.params())
            if isinstance(l, E.BatchNorm):
                ws += [l.moving_mean, l.moving_var]
            for p in ws:
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out

    def get_weights(self):
        self._ensure_device()
        torch.cuda.synchronize()
        return [p.numpy() for p in self.weight_list()]

    def set_weights(self, weights):
        self._ensure_device()
        wl = self.weight_list()
        assert len(wl) == len(weights), "%s: expected %d arrays, got %d" % (self.name, len(wl), len(weights))
        for p, w in zip(wl, weights):
            w = np.asarray(w, np.float32)
            assert tuple(w.shape) == p.shape, (p.name, w.shape, p.shape)
            p.data.copy_(torch.from_numpy(np.ascontiguousarray(w)))
        self._scope.arena.version += 1
        self._scope.state.version += 1          # BatchNorm moving statistics may have changed too

    def named_weights(self):
        return {p.name: p.numpy() for p in self.weight_list()}

    def save_weights(self, path):
        """the reference writes Keras HDF5 (dafnet_executor.py:286-301); h5py is not installed, so the
        same per-layer arrays go into an .npz keyed by weight name, in Keras order."""
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        ws = self.get_weights()
        names = [p.name for p in self.weight_list()]
        np.savez(path if path.endswith(".npz") else path + ".npz", __order__=np.array(names), **dict(zip(names, ws)))

    def load_weights(self, path):
        f = path if path.endswith(".npz") else path + ".npz"
        z = np.load(f, allow_pickle=False)
        self.set_weights([z[n] for n in z["__order__"]])

    def count_params(self):
        return int(sum(p.size for p in self.params()))

    def summary(self, print_fn=print):
        print_fn("Model: %s" % self.name)
        for l in self.layers:
            print_fn("  %-28s %s" % (l.name, " ".join("%s%s" % (p.name.split("/")[-1], p.shape) for p in l.params())))
        print_fn("Total params: %d" % self.count_params())

    def predict(self, x, batch_size=32):
        """Keras ``predict``: numpy in, numpy out, inference phase (moving BN statistics), batches of 32."""
        self._ensure_device()
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        n = xs[0].shape[0]
        outs = None
        for s in range(0, n, batch_size):
            vs = [E.Var(torch.as_tensor(np.ascontiguousarray(a[s:s + batch_size], dtype=np.float32)).cuda()) for a in xs]
            r = self._forward(E.Ctx(None, training=False), *vs)
            r = list(r) if isinstance(r, (list, tuple)) else [r]
            r = [o.data.float().cpu().numpy() for o in r]
            if outs is None:
                outs = [[] for _ in r]
            for acc, o in zip(outs, r):
                acc.append(o)
        outs = [np.concatenate(o, 0) for o in outs]
        return outs[0] if len(outs) == 1 else outs

    def predict_device(self, *tensors):
        """same as predict but device tensors in / out (no host round trip)."""
        self._ensure_device()
        r = self._forward(E.Ctx(None, training=False), *[E.Var(t) for t in tensors])
        if isinstance(r, (list, tuple)):
            return [o.data for o in r]
        return r.data

"""numpy stand-ins for the handful of TensorFlow-1.x / Keras-2.1 entry points that the reference's layer and
loss code calls, so that the reference's OWN Python source (layers/interpolate_spline.py, layers/stn_spline.py,
layers/film.py, layers/spade.py, layers/rounding.py, layers/spectralnorm.py, costs.py, utils/sdnet_utils.py,
model_executors/base_executor.py, utils/data_utils.py) can be imported unmodified from /root/reference and
executed eagerly to produce golden vectors (tests/golden/make_golden.py).

TF 1.4 / Keras 2.1.6 cannot be installed in this image (py3.12, no wheels, no network).  Every function here
implements the documented semantics of the TF/Keras op of the same name on numpy arrays -- the ALGORITHMS under
test stay the reference's own code.  This file is test infrastructure only.
"""
import contextlib
import sys
import types

import numpy as np

from tests.golden import keras_graph as KG


class T(np.ndarray):
    """ndarray with the little bit of tf.Tensor surface the reference touches"""

    def get_shape(self):
        return _Shape(self.shape)

    def set_shape(self, shape):
        return None


class _Dim(int):
    @property
    def value(self):
        return int(self)


class _Shape(tuple):
    def __new__(cls, s):
        return super().__new__(cls, [_Dim(d) for d in s])

    def as_list(self):
        return [int(d) for d in self]


def t(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    return a.view(T)


def _axis(axis=None, reduction_indices=None):
    a = axis if axis is not None else reduction_indices
    if isinstance(a, (list, tuple)):
        return tuple(int(i) for i in a)
    return a


# ------------------------------------------------------------------------------------------- tensorflow
def _make_tf():
    tf = types.ModuleType("tensorflow")
    tf.float32 = np.float32
    tf.float = np.float32
    tf.float64 = np.float64
    tf.reduce_sum = lambda x, axis=None, reduction_indices=None, name=None, keepdims=False: t(
        np.sum(np.asarray(x), axis=_axis(axis, reduction_indices), keepdims=keepdims))
    tf.reduce_mean = lambda x, axis=None, reduction_indices=None, name=None, keepdims=False: t(
        np.mean(np.asarray(x), axis=_axis(axis, reduction_indices), keepdims=keepdims))
    tf.reshape = lambda x, shape, name=None: t(np.reshape(np.asarray(x), tuple(int(s) for s in shape)))
    tf.multiply = lambda a, b, name=None: t(np.asarray(a) * np.asarray(b))
    tf.log = lambda x: t(np.log(np.asarray(x)))
    tf.to_float = lambda x: t(np.asarray(x))          # goldens are produced in the caller's dtype
    tf.cast = lambda x, dtype=None, name=None: t(np.asarray(x)) if dtype in (np.float32, "float32") else t(np.asarray(x), dtype)
    tf.tile = lambda x, reps: t(np.tile(np.asarray(x), tuple(int(r) for r in reps)))
    tf.expand_dims = lambda x, axis=-1: t(np.expand_dims(np.asarray(x), axis))
    tf.reverse = lambda x, axis: t(np.flip(np.asarray(x), axis=tuple(axis)))
    tf.map_fn = lambda fn, elems: t(np.stack([np.asarray(fn(t(e))) for e in np.asarray(elems)], 0))
    tf.maximum = lambda a, b: t(np.maximum(np.asarray(a), np.asarray(b)))
    tf.concat = lambda xs, axis: t(np.concatenate([np.asarray(x) for x in xs], axis))

    nn = types.ModuleType("tensorflow.nn")

    def softmax(x, axis=-1, dim=None):          # TF 1.4 spells the axis `dim` (model_components/balancer.py:26)
        axis = axis if dim is None else dim
        x = np.asarray(x)
        e = np.exp(x - x.max(axis=axis, keepdims=True))
        return t(e / e.sum(axis=axis, keepdims=True))
    nn.softmax = softmax
    tf.nn = nn

    image = types.ModuleType("tensorflow.image")

    def resize_nearest_neighbor(x, size, align_corners=False):
        x = np.asarray(x)
        H, W = x.shape[1], x.shape[2]
        ho, wo = int(size[0]), int(size[1])
        iy = np.minimum(np.floor(np.arange(ho) * (H / ho)).astype(int), H - 1)
        ix = np.minimum(np.floor(np.arange(wo) * (W / wo)).astype(int), W - 1)
        return t(x[:, iy][:, :, ix])
    image.resize_nearest_neighbor = resize_nearest_neighbor
    tf.image = image

    # tf.contrib.resampler.resampler: source not vendored in the reference; documented kernel
    contrib = types.ModuleType("tensorflow.contrib")
    resampler = types.ModuleType("tensorflow.contrib.resampler")

    def resampler_fn(data, warp):
        data = np.asarray(data)
        warp = np.asarray(warp)
        B, H, W, C = data.shape
        out = np.zeros(warp.shape[:-1] + (C,), data.dtype)
        flat = warp.reshape(B, -1, 2)
        o = out.reshape(B, -1, C)
        for b in range(B):
            x, y = flat[b, :, 0], flat[b, :, 1]
            ok = (x > -1) & (y > -1) & (x < W) & (y < H)
            fx, fy = np.floor(x), np.floor(y)
            cx, cy = fx + 1, fy + 1
            dx, dy = cx - x, cy - y

            def at(ix, iy):
                inside = (ix >= 0) & (ix <= W - 1) & (iy >= 0) & (iy <= H - 1)
                v = data[b, np.clip(iy, 0, H - 1).astype(int), np.clip(ix, 0, W - 1).astype(int)]
                return v * inside[:, None]
            val = (dx * dy)[:, None] * at(fx, fy) + ((1 - dx) * (1 - dy))[:, None] * at(cx, cy) \
                + (dx * (1 - dy))[:, None] * at(fx, cy) + ((1 - dx) * dy)[:, None] * at(cx, fy)
            o[b] = val * ok[:, None]
        return t(out)
    resampler.resampler = resampler_fn
    contrib.resampler = resampler
    eager = types.ModuleType("tensorflow.contrib.eager")
    eager.defun = lambda f: f
    contrib.eager = eager
    tf.contrib = contrib

    class _Graph:
        @contextlib.contextmanager
        def gradient_override_map(self, m):
            yield
    tf.get_default_graph = lambda: _Graph()
    tf.RegisterGradient = lambda name: (lambda fn: fn)
    tf.py_func = lambda func, inp, Tout, stateful=True, name=None: [t(func(*[np.asarray(i) for i in inp]))]

    # tensorflow.python.framework / ops
    python = types.ModuleType("tensorflow.python")
    framework = types.ModuleType("tensorflow.python.framework")
    ops = types.ModuleType("tensorflow.python.framework.ops")

    @contextlib.contextmanager
    def name_scope(name=None, default_name=None, values=None):
        yield name or default_name
    ops.name_scope = name_scope
    ops.convert_to_tensor = lambda x, name=None, dtype=None: t(x)
    tensor_shape = types.ModuleType("tensorflow.python.framework.tensor_shape")
    tensor_shape.dimension_value = lambda d: int(d)
    framework.ops = ops
    framework.tensor_shape = tensor_shape
    pops = types.ModuleType("tensorflow.python.ops")
    array_ops = types.ModuleType("tensorflow.python.ops.array_ops")
    array_ops.concat = lambda xs, axis: t(np.concatenate([np.asarray(x) for x in xs], axis))
    array_ops.expand_dims = lambda x, axis: t(np.expand_dims(np.asarray(x), axis))
    array_ops.zeros = lambda shape, dtype=np.float32: t(np.zeros(tuple(int(s) for s in shape), dtype))
    array_ops.transpose = lambda x, perm=None: t(np.transpose(np.asarray(x), perm))
    array_ops.ones_like = lambda x, dtype=None: t(np.ones_like(np.asarray(x), dtype=dtype))
    array_ops.unstack = lambda x, num=None: [int(v) for v in np.asarray(x)]
    array_ops.shape = lambda x: np.asarray(np.asarray(x).shape)
    array_ops.matrix_diag_part = lambda x: t(np.diagonal(np.asarray(x), axis1=-2, axis2=-1).copy())
    linalg_ops = types.ModuleType("tensorflow.python.ops.linalg_ops")
    linalg_ops.matrix_solve = lambda a, b: t(np.linalg.solve(np.asarray(a), np.asarray(b)))   # LU, partial pivoting
    linalg_ops.eye = lambda n, dtype=np.float32: t(np.eye(int(n), dtype=dtype))
    math_ops = types.ModuleType("tensorflow.python.ops.math_ops")
    math_ops.reduce_sum = tf.reduce_sum
    math_ops.square = lambda x: t(np.square(np.asarray(x)))
    math_ops.maximum = tf.maximum
    math_ops.log = tf.log
    math_ops.sqrt = lambda x: t(np.sqrt(np.asarray(x)))
    math_ops.pow = lambda x, p: t(np.power(np.asarray(x), p))

    def matmul(a, b, adjoint_b=False, transpose_b=False):
        b = np.asarray(b)
        if adjoint_b or transpose_b:
            b = np.swapaxes(b, -1, -2)
        return t(np.matmul(np.asarray(a), b))
    math_ops.matmul = matmul
    tf.matmul = matmul
    pops.array_ops, pops.linalg_ops, pops.math_ops = array_ops, linalg_ops, math_ops
    python.framework, python.ops = framework, pops
    tf.python = python
    mods = {"tensorflow": tf, "tensorflow.nn": nn, "tensorflow.image": image, "tensorflow.contrib": contrib,
            "tensorflow.contrib.resampler": resampler, "tensorflow.contrib.eager": eager, "tensorflow.python": python,
            "tensorflow.python.framework": framework, "tensorflow.python.framework.ops": ops,
            "tensorflow.python.framework.tensor_shape": tensor_shape, "tensorflow.python.ops": pops,
            "tensorflow.python.ops.array_ops": array_ops, "tensorflow.python.ops.linalg_ops": linalg_ops,
            "tensorflow.python.ops.math_ops": math_ops}
    return mods


# ------------------------------------------------------------------------------------------- keras
class _Layer(object):
    """keras.engine.topology.Layer for the reference's own layer classes: eager on arrays, a graph node on the symbolic
    tensors of tests/golden/keras_graph.py (the builders' functional-API use)"""

    def __init__(self, **kwargs):
        self.name = kwargs.get("name")
        self.built = False
        self.ctor = KG.STATE["ctor"]
        KG.STATE["ctor"] += 1

    def build(self, input_shape):
        self.built = True

    def __call__(self, x, **kwargs):
        if KG._has_sym(x):
            return KG.sym_call(self, x)
        if not self.built:
            self.build(None)
        # arrays coming out of the numpy graph layers get the little bit of tf.Tensor surface back
        x = [t(v) if isinstance(v, np.ndarray) else v for v in x] if isinstance(x, (list, tuple)) else (
            t(x) if isinstance(x, np.ndarray) else x)
        # Keras executes `call` ONCE, while the graph is built; here it runs at every evaluation.  Graph-construction side
        # effects on numpy's global generator (layers/rounding.py:22 draws a random op name) must therefore not leak into
        # the data path of a training step (utils.data_utils.sample draws from the same generator).
        state = np.random.get_state()
        try:
            return self.call(x, **kwargs)
        finally:
            np.random.set_state(state)


class _Dummy(object):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise NotImplementedError("graph-building Keras objects are not part of the golden shim")


def _make_keras(rng_holder):
    keras = types.ModuleType("keras")
    K = types.ModuleType("keras.backend")
    K.sum = lambda x, axis=None, keepdims=False: t(np.sum(np.asarray(x), axis=_axis(axis), keepdims=keepdims))
    K.mean = lambda x, axis=None, keepdims=False: t(np.mean(np.asarray(x), axis=_axis(axis), keepdims=keepdims))
    K.abs = lambda x: t(np.abs(np.asarray(x)))
    K.square = lambda x: t(np.square(np.asarray(x)))
    K.sqrt = lambda x: t(np.sqrt(np.asarray(x)))
    K.exp = lambda x: t(np.exp(np.asarray(x)))
    K.shape = lambda x: tuple(int(s) for s in np.asarray(x).shape)
    K.int_shape = lambda x: tuple(int(s) for s in np.asarray(x).shape)
    K.reshape = lambda x, shape: t(np.reshape(np.asarray(x), tuple(int(s) for s in shape)))
    K.tile = lambda x, n: t(np.tile(np.asarray(x), tuple(int(r) for r in n)))
    K.expand_dims = lambda x, axis=-1: t(np.expand_dims(np.asarray(x), axis))
    K.transpose = lambda x: t(np.transpose(np.asarray(x)))
    K.dot = lambda a, b: t(np.dot(np.asarray(a), np.asarray(b)))
    K.stop_gradient = lambda x: x
    K.variable = lambda v, dtype=None, name=None: t(np.asarray(v, dtype=np.float64))
    K.epsilon = lambda: 1e-7
    K.random_normal = lambda shape, mean=0.0, stddev=1.0: t(rng_holder["rng"].normal(mean, stddev, tuple(shape)))
    keras.backend = K
    engine = types.ModuleType("keras.engine")
    engine.Layer = _Layer
    topology = types.ModuleType("keras.engine.topology")
    topology.Layer = _Layer
    engine.topology = topology
    layers = types.ModuleType("keras.layers")
    for n in ("Concatenate", "MaxPooling2D", "Conv2D", "Flatten", "Dense", "Reshape", "LeakyReLU", "Lambda", "Add",
              "Activation", "UpSampling2D", "BatchNormalization", "Input", "Maximum", "Multiply"):
        setattr(layers, n, getattr(KG, n))          # define-then-run numpy layers (tests/golden/keras_graph.py)
    keras.layers = layers
    keras.Input = KG.Input
    keras.Model = KG.Model
    optimizers = types.ModuleType("keras.optimizers")
    optimizers.Adam = _Dummy
    keras.optimizers = optimizers
    regularizers = types.ModuleType("keras.regularizers")
    regularizers.Regularizer = object
    keras.regularizers = regularizers
    pre = types.ModuleType("keras.preprocessing")
    pimg = types.ModuleType("keras.preprocessing.image")
    pimg.ImageDataGenerator = _Dummy
    pre.image = pimg
    keras.preprocessing = pre
    cbs = types.ModuleType("keras.callbacks")
    for n in ("Callback", "CSVLogger", "EarlyStopping"):
        setattr(cbs, n, _Dummy)
    keras.callbacks = cbs
    utils = types.ModuleType("keras.utils")
    utils.Progbar = _Dummy
    keras.utils = utils
    return {"keras": keras, "keras.backend": K, "keras.engine": engine, "keras.engine.topology": topology,
            "keras.layers": layers, "keras.regularizers": regularizers, "keras.preprocessing": pre,
            "keras.preprocessing.image": pimg, "keras.callbacks": cbs, "keras.utils": utils,
            "keras.optimizers": optimizers}


RNG = {"rng": np.random.RandomState(0)}


@contextlib.contextmanager
def installed(reference_root="/root/reference"):
    """Temporarily install the shim modules and put the reference on sys.path."""
    mods = {}
    mods.update(_make_tf())
    mods.update(_make_keras(RNG))
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    sys.path.insert(0, reference_root)
    try:
        yield
    finally:
        sys.path.remove(reference_root)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k.split(".")[0] in ("layers", "costs", "utils", "model_executors", "loaders",
                                                                 "callbacks", "model_components", "models")]:
            # the reference's top-level package names must not leak into the test process
            m = sys.modules[k]
            f = getattr(m, "__file__", "") or ""
            if f.startswith(reference_root):
                sys.modules.pop(k, None)

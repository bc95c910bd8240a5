"""Define-then-run numpy stand-in for the slice of the Keras 2.1.6 functional API that the reference's BUILDERS use
(model_components/*.py, models/unet.py, models/discriminator.py, layers/stn_spline.build_locnet): Input, Model (also
called as a layer), Conv2D, Dense, BatchNormalization (inference phase), Activation, LeakyReLU, MaxPooling2D,
UpSampling2D, Concatenate, Add, Maximum, Flatten, Reshape, Lambda.

tests/golden/make_golden.py runs the reference's own builder code on top of it, so the golden outputs pin the WIRING of
every component (which layer feeds which, shared layers, Keras weight order) -- the layer arithmetic itself is the
documented Keras semantics restated here in float64 numpy (Conv2D SAME/VALID cross-correlation, BatchNormalization with
epsilon 1e-3 on the moving statistics, LeakyReLU slope 0.3 by default, ...).  Test infrastructure only.

Weights: every weighted layer draws small integers from a generator keyed by its CONSTRUCTION index and scales them
(w = float32(offset + k * scale)), whatever initializer the builder asked for -- zero-initialised heads would hide wiring
mistakes.  A model's weight list is "weighted layers reachable from its outputs, by construction index; per layer kernel,
bias / gamma, beta, moving_mean, moving_var" (for the sequentially built components this is Keras' get_weights order).
"""
import numpy as np

STATE = {"ctor": 0, "seed": 1234}


def reset(seed=1234):
    STATE["ctor"] = 0
    STATE["seed"] = seed


def _wrap(a):
    from tests.golden.tf_shim import t     # the tf.Tensor-like ndarray view the reference's loss functions expect
    return t(a)


class Sym(object):
    """output `index` of `layer` applied to `inputs` (a Sym or a list of Syms); layer None = placeholder"""

    def __init__(self, layer, inputs, index=0):
        self.layer, self.inputs, self.index = layer, inputs, index


def _has_sym(x):
    return isinstance(x, Sym) or (isinstance(x, (list, tuple)) and any(isinstance(v, Sym) for v in x))


def evaluate(sym, feeds, cache):
    if id(sym) in cache:
        return cache[id(sym)]
    if sym.layer is None:
        val = feeds[id(sym)]
    else:
        key = ("call", id(sym.layer), id(sym.inputs) if isinstance(sym.inputs, Sym) else tuple(id(v) for v in sym.inputs))
        if key not in cache:
            if isinstance(sym.inputs, Sym):
                xin = evaluate(sym.inputs, feeds, cache)
            else:
                xin = [evaluate(v, feeds, cache) for v in sym.inputs]
            cache[key] = sym.layer(xin)
        out = cache[key]
        val = out[sym.index] if isinstance(out, (list, tuple)) else out
    cache[id(sym)] = val
    return val


def sym_call(layer, x):
    """what tf_shim._Layer.__call__ and the layers below return for symbolic inputs"""
    s = Sym(layer, list(x) if isinstance(x, (list, tuple)) else x)
    layer.output = s
    return s


class Layer(object):
    def __init__(self, name=None, **kwargs):
        self.name = name
        self.built = False
        self.ctor = STATE["ctor"]
        STATE["ctor"] += 1
        self.k, self.scale_offset, self.w = [], [], []       # integer draws, (scale, offset), float32 values

    def _draw(self, shape, scale, offset=0.0):
        rs = np.random.RandomState(STATE["seed"] + 7919 * self.ctor + len(self.k))
        k = rs.randint(-32, 33, size=shape).astype(np.int8)
        self.k.append(k)
        self.scale_offset.append(np.array([scale, offset], np.float64))
        self.w.append((offset + k.astype(np.float64) * scale).astype(np.float32))
        return self.w[-1].astype(np.float64)

    def build(self, shape):
        pass

    def __call__(self, x, **kwargs):
        if _has_sym(x):
            return sym_call(self, x)
        if not self.built:
            self.build(x[0].shape if isinstance(x, (list, tuple)) else x.shape)
            self.built = True
        return self.call(x)


def _act(x, name):
    if name in (None, "linear"):
        return x
    if name == "relu":
        return np.maximum(x, 0.0)
    if name == "tanh":
        return np.tanh(x)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    if name == "softmax":
        e = np.exp(x - x.max(-1, keepdims=True))
        return e / e.sum(-1, keepdims=True)
    raise ValueError(name)


class Input(Sym):
    def __init__(self, shape=None, **kwargs):
        Sym.__init__(self, None, None)
        self.shape_ = tuple(shape)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, name=None, use_bias=True,
                 **kwargs):
        Layer.__init__(self, name)
        self.f, self.ks = int(filters), int(kernel_size if np.isscalar(kernel_size) else kernel_size[0])
        self.s = int(strides if np.isscalar(strides) else strides[0])
        self.padding, self.activation, self.use_bias = padding, activation, use_bias
        self.kernel_regularizer = kwargs.get("kernel_regularizer")

    def build(self, shape):
        cin = shape[-1]
        fan_in = self.ks * self.ks * cin
        self.kernel = self._draw((self.ks, self.ks, cin, self.f), 4.9 / 64.0 / np.sqrt(fan_in))
        self.bias = self._draw((self.f,), 0.1 / 64.0) if self.use_bias else 0.0

    def assign(self, ws):
        self.kernel = ws[0].astype(np.float64)
        if self.use_bias:
            self.bias = ws[1].astype(np.float64)

    def call(self, x):
        x = np.asarray(x, np.float64)
        k, s = self.ks, self.s
        if self.padding == "same":      # TensorFlow SAME: out = ceil(in / s), the extra pixel goes to the bottom / right
            pads = []
            for n in x.shape[1:3]:
                out = -(-n // s)
                tot = max((out - 1) * s + k - n, 0)
                pads.append((tot // 2, tot - tot // 2))
            x = np.pad(x, ((0, 0), pads[0], pads[1], (0, 0)))
        win = np.lib.stride_tricks.sliding_window_view(x, (k, k), axis=(1, 2))[:, ::s, ::s]    # N, Ho, Wo, C, k, k
        y = np.einsum("nhwcij,ijco->nhwo", win, self.kernel, optimize=True) + self.bias
        return _act(y, self.activation)


class Dense(Layer):
    def __init__(self, units, activation=None, name=None, **kwargs):
        Layer.__init__(self, name)
        self.units, self.activation = int(units), activation

    def build(self, shape):
        self.kernel = self._draw((shape[-1], self.units), 4.9 / 64.0 / np.sqrt(shape[-1]))
        self.bias = self._draw((self.units,), 0.1 / 64.0)

    def assign(self, ws):
        self.kernel, self.bias = ws[0].astype(np.float64), ws[1].astype(np.float64)

    def call(self, x):
        return _act(np.asarray(x, np.float64) @ self.kernel + self.bias, self.activation)


class BatchNormalization(Layer):
    """inference phase: gamma * (x - moving_mean) / sqrt(moving_var + 1e-3) + beta"""

    def build(self, shape):
        c = shape[-1]
        self.gamma = self._draw((c,), 1.0 / 128.0, 1.0)
        self.beta = self._draw((c,), 1.0 / 256.0)
        self.mm = self._draw((c,), 1.0 / 256.0)
        self.mv = self._draw((c,), 1.0 / 128.0, 1.0)

    def call(self, x):
        return self.gamma * (np.asarray(x, np.float64) - self.mm) / np.sqrt(self.mv + 1e-3) + self.beta


class Activation(Layer):
    def __init__(self, activation, name=None, **kwargs):
        Layer.__init__(self, name)
        self.activation = activation

    def call(self, x):
        return _act(np.asarray(x, np.float64), self.activation)


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kwargs):
        Layer.__init__(self, kwargs.get("name"))
        self.alpha = alpha

    def call(self, x):
        x = np.asarray(x, np.float64)
        return np.where(x > 0, x, self.alpha * x)


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), **kwargs):
        Layer.__init__(self, kwargs.get("name"))
        assert tuple(pool_size) == (2, 2)

    def call(self, x):
        x = np.asarray(x, np.float64)
        n, h, w, c = x.shape
        x = x[:, :h // 2 * 2, :w // 2 * 2]
        return x.reshape(n, h // 2, 2, w // 2, 2, c).max(axis=(2, 4))


class UpSampling2D(Layer):
    def __init__(self, size=2, **kwargs):
        Layer.__init__(self, kwargs.get("name"))
        self.size = int(size if np.isscalar(size) else size[0])

    def call(self, x):
        return np.repeat(np.repeat(np.asarray(x, np.float64), self.size, 1), self.size, 2)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        Layer.__init__(self, kwargs.get("name"))
        self.axis = axis

    def call(self, xs):
        return np.concatenate([np.asarray(v, np.float64) for v in xs], self.axis)


class Add(Layer):
    def call(self, xs):
        return sum(np.asarray(v, np.float64) for v in xs)


class Multiply(Layer):
    def call(self, xs):
        out = np.asarray(xs[0], np.float64)
        for v in xs[1:]:
            out = out * np.asarray(v, np.float64)
        return out


class Maximum(Layer):
    def call(self, xs):
        out = np.asarray(xs[0], np.float64)
        for v in xs[1:]:
            out = np.maximum(out, np.asarray(v, np.float64))
        return out


class Flatten(Layer):
    def call(self, x):
        x = np.asarray(x, np.float64)
        return x.reshape(x.shape[0], -1)


class Reshape(Layer):
    def __init__(self, target_shape, **kwargs):
        Layer.__init__(self, kwargs.get("name"))
        self.target = tuple(int(v) for v in target_shape)

    def call(self, x):
        x = np.asarray(x, np.float64)
        return x.reshape((x.shape[0],) + self.target)


def _freeze(fn):
    """Keras runs a Lambda's function while the graph is being BUILT, so a closure over a loop variable
    (models/dafnet.py:359 `lambda x: x[..., j:j+1]` inside a comprehension) sees that iteration's value.  Here the
    function only runs at evaluation time: snapshot the closure cells when the layer is applied."""
    if getattr(fn, "__closure__", None) is None:
        return fn
    import types
    cells = tuple(types.CellType(c.cell_contents) for c in fn.__closure__)
    return types.FunctionType(fn.__code__, fn.__globals__, fn.__name__, fn.__defaults__, cells)


class Lambda(Layer):
    """`arguments` may hold symbolic tensors (layers/spade.py:30 passes the tensor to resize to): they become extra inputs
    of the node and are handed to the function by keyword when it runs"""

    def __init__(self, function, name=None, arguments=None, **kwargs):
        Layer.__init__(self, name)
        self.fn = function
        self.arguments = dict(arguments or {})
        self.sym_keys = [k for k, v in self.arguments.items() if isinstance(v, Sym)]

    def __call__(self, x, **kwargs):
        if _has_sym(x):
            self.fn = _freeze(self.fn)
            if self.sym_keys:
                assert isinstance(x, Sym)
                return sym_call(self, [x] + [self.arguments[k] for k in self.sym_keys])
        return Layer.__call__(self, x, **kwargs)

    def call(self, x):
        if self.sym_keys:
            kw = dict(self.arguments)
            kw.update({k: _wrap(v) for k, v in zip(self.sym_keys, x[1:])})
            return np.asarray(self.fn(_wrap(x[0]), **kw), np.float64)
        return np.asarray(self.fn(x, **self.arguments), np.float64)


class InstanceNormalization(Layer):
    """keras_contrib 2.0.8, axis=None, no scale / centre: (x - mean) / (std + epsilon) over all of H, W, C per sample"""

    def __init__(self, axis=None, epsilon=1e-3, center=True, scale=True, **kwargs):
        Layer.__init__(self, kwargs.get("name"))
        assert axis is None and not center and not scale
        self.epsilon = epsilon

    def call(self, x):
        x = np.asarray(x, np.float64)
        ax = tuple(range(1, x.ndim))
        return (x - x.mean(ax, keepdims=True)) / (x.std(ax, keepdims=True) + self.epsilon)


class Model(Layer):
    def __init__(self, inputs=None, outputs=None, name=None, input=None, output=None, **kwargs):
        Layer.__init__(self, name)
        inputs = inputs if inputs is not None else input
        outputs = outputs if outputs is not None else output
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.single_out = not isinstance(outputs, (list, tuple))
        self.outputs = [outputs] if self.single_out else list(outputs)
        self.built = True

    def __call__(self, x, **kwargs):
        if _has_sym(x):
            xs = list(x) if isinstance(x, (list, tuple)) else [x]
            syms = [Sym(self, xs, i) for i in range(len(self.outputs))]
            self.output = syms[0] if self.single_out else syms
            return self.output
        return self.call(x)

    def call(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        assert len(xs) == len(self.inputs), (self.name, len(xs), len(self.inputs))
        feeds = {id(s): np.asarray(v, np.float64) for s, v in zip(self.inputs, xs)}
        cache = {}
        return [evaluate(o, feeds, cache) for o in self.outputs]     # always a list; Sym.index picks

    def predict(self, x):
        out = self.call(x)
        return out[0] if self.single_out else out

    # ---- the bits of the Keras Model surface the builders touch
    def summary(self, print_fn=None, **kwargs):
        return None

    def compile(self, optimizer=None, loss=None, loss_weights=None, **kwargs):
        self.loss, self.loss_weights = loss, loss_weights

    def output_names(self):
        """Keras names an output after the layer that produced it (duplicates allowed); a loss / weight dictionary is
        looked up by that name, a list by position (keras/engine/training.py, compile)"""
        return [o.layer.name for o in self.outputs]

    def loss_values(self, x, targets):
        """weight_i * mean(loss_i(y_true_i, y_pred_i)) per output, as Keras' total loss sums them (unit sample weights)"""
        outs = self.call(x)
        names = self.output_names()
        vals = []
        for i, (name, yp, yt) in enumerate(zip(names, outs, targets)):
            fn = self.loss[name] if isinstance(self.loss, dict) else (self.loss[i] if isinstance(self.loss, (list, tuple)) else self.loss)
            w = 1.0 if self.loss_weights is None else (self.loss_weights[name] if isinstance(self.loss_weights, dict)
                                                       else self.loss_weights[i])
            yp, yt = np.asarray(yp, np.float64), np.asarray(yt, np.float64)
            if fn == "mse":
                v = np.mean(np.square(yp - yt), axis=-1)
            elif fn == "mae":
                v = np.mean(np.abs(yp - yt), axis=-1)
            else:
                v = np.asarray(fn(_wrap(yt), _wrap(yp)), np.float64)
            vals.append(float(w) * float(np.mean(v)))
        return vals

    @property
    def output_shape(self):
        """static shapes by running the graph once on a zero batch of one sample"""
        outs = self.call([np.zeros((1,) + s.shape_) for s in self.inputs])
        shp = [(None,) + tuple(np.shape(o)[1:]) for o in outs]
        return shp[0] if self.single_out else shp

    def get_output_shape_at(self, index):
        return self.output_shape

    def fit(self, x, y, **kwargs):
        """no training here (the stand-in has no autodiff): record what the executor feeds and return the per-name mean
        losses Keras would log, so that the reference's executor code runs unmodified"""
        class _H(object):
            pass
        xs = [np.asarray(v) for v in (x if isinstance(x, (list, tuple)) else [x])]
        ys = [np.asarray(v) for v in (y if isinstance(y, (list, tuple)) else [y])]
        if not hasattr(self, "fit_calls"):
            self.fit_calls = []
        self.fit_calls.append((xs, ys))
        STATE.setdefault("fit_log", []).append(str(self.name))
        vals = self.loss_values(xs, [v.reshape(v.shape[0], -1) if v.ndim == 1 else v for v in ys])
        h = _H()
        h.history = {"loss": [float(np.sum(vals))]}
        for name, v in zip(self.output_names(), vals):
            key = "%s_loss" % name                    # unnamed sub-models (Enc_Modality_mu): Keras would auto-name them
            h.history.setdefault(key, [0.0])
            h.history[key][0] += v
        return h

    def load_weights(self, path):
        raise IOError("the golden run starts from freshly drawn weights (%s)" % path)

    def _walk(self):
        seen, order = set(), []

        def visit(s):
            if id(s) in seen or s.layer is None:
                return
            seen.add(id(s))
            for v in ([s.inputs] if isinstance(s.inputs, Sym) else s.inputs):
                visit(v)
            lay = s.layer
            if isinstance(lay, Model):
                for l in lay._walk():
                    if l not in order:
                        order.append(l)
            if lay not in order:
                order.append(lay)
        for o in self.outputs:
            visit(o)
        return order

    @property
    def layers(self):
        return sorted(self._walk(), key=lambda l: l.ctor)

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError("No such layer: " + str(name))

    def weighted_layers(self):
        return [l for l in self.layers if getattr(l, "w", None)]

    def get_weights(self):
        return [w for l in self.weighted_layers() for w in l.w]

    def regularized_layers(self):
        """layers whose kernel carries a regulariser (Keras adds regulariser(kernel) to the model's total loss)"""
        return [l for l in self.weighted_layers() if getattr(l, "kernel_regularizer", None) is not None]

    def set_weights(self, ws):
        """replace the drawn weights (Conv2D / Dense layers), in get_weights order"""
        pos = 0
        for l in self.weighted_layers():
            n = len(l.w)
            l.w = [np.asarray(w, np.float32) for w in ws[pos:pos + n]]
            l.assign(l.w)
            pos += n
        assert pos == len(ws)

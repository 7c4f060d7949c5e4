"""Core of the Keras-2.2.4 stand-in (test infrastructure; see ../README.md).  torch-CPU autograd underneath.

Restated from the published Keras 2.2.4 sources (file names below are paths inside the `keras` 2.2.4 wheel):
engine/base_layer.py (Layer naming, add_weight, __call__/build protocol), engine/network.py (graph walk, get_layer,
get_weights/set_weights order = layer order), engine/training.py + training_arrays.py (fit: one shuffled index array
per epoch, batches, History keys, total loss = sum of the per-output losses), engine/training_utils.py
(weighted_masked_objective), optimizers.py (Adam), initializers.py (glorot_uniform), backend/tensorflow_backend.py
(dot, bias_add, concatenate, batch_dot), backend/common.py (floatx float32, epsilon 1e-7).
"""
import os
import re

import numpy as np
import torch

_FLOATX = os.environ.get("KERAS_SHIM_FLOATX", "float32")        # backend/common.py: _FLOATX = 'float32'
_EPSILON = 1e-7                                                  # backend/common.py: _EPSILON = 1e-7
_UIDS = {}
_INIT_RNG = np.random.RandomState(12345)    # TF draws initial weights from its own generator, never from np.random


def floatx():
    return _FLOATX


def epsilon():
    return _EPSILON


def _tdtype():
    return getattr(torch, _FLOATX)


def get_uid(prefix=""):                                          # backend get_uid: per-graph counters from 1
    _UIDS[prefix] = _UIDS.get(prefix, 0) + 1
    return _UIDS[prefix]


def reset_uids():
    _UIDS.clear()


def _to_snake_case(name):                                        # engine/base_layer.py _to_snake_case
    intermediate = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    insecure = re.sub("([a-z])([A-Z])", r"\1_\2", intermediate).lower()
    return insecure if insecure[0] != "_" else "private" + insecure


# ------------------------------------------------------------------ activations / initializers
class _Activations:
    @staticmethod
    def relu(x):
        return torch.relu(x)

    @staticmethod
    def linear(x):
        return x

    def get(self, identifier):                                   # activations.get: None -> linear
        if identifier is None:
            return self.linear
        if callable(identifier):
            return identifier
        return getattr(self, identifier)


activations = _Activations()


def _initial_value(initializer, shape):
    if initializer == "zeros":
        return np.zeros(shape)
    if initializer == "glorot_uniform":                          # initializers.py VarianceScaling(1, fan_avg, uniform)
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
        limit = np.sqrt(3.0 * 1.0 / max(1.0, (fan_in + fan_out) / 2.0))
        return _INIT_RNG.uniform(-limit, limit, shape)
    raise ValueError(f"initializer {initializer!r} not in the shim")


# ------------------------------------------------------------------ symbolic tensors and layers
class KTensor:
    """A symbolic tensor: static shape (None, ...) and the (layer, call index, output index) that produces it."""

    def __init__(self, shape, layer, node_index, tensor_index):
        self._keras_shape = tuple(shape)
        self._history = (layer, node_index, tensor_index)


def _shapes_of(x):
    return [t._keras_shape for t in x] if isinstance(x, (list, tuple)) else x._keras_shape


class Layer:
    def __init__(self, **kwargs):
        name = kwargs.pop("name", None)
        assert not kwargs or set(kwargs) <= {"trainable", "dtype", "input_shape"}, kwargs
        if not name:
            prefix = _to_snake_case(self.__class__.__name__)
            name = prefix + "_" + str(get_uid(prefix))
        self.name = name
        self.built = False
        self.trainable_weights = []
        self._weight_names = []
        self._inbound = []                                       # one entry per call: the symbolic inputs

    def add_weight(self, name, shape, initializer=None, trainable=True, **unused):
        w = torch.tensor(_initial_value(initializer, tuple(shape)), dtype=_tdtype(), requires_grad=bool(trainable))
        self.trainable_weights.append(w)
        self._weight_names.append(name)
        return w

    def build(self, input_shape):
        self.built = True

    def call(self, inputs):
        return inputs

    def compute_output_shape(self, input_shape):
        return input_shape

    def __call__(self, inputs, **kwargs):
        shapes = _shapes_of(inputs)
        if not self.built:
            self.build(shapes)
            self.built = True
        out_shape = self.compute_output_shape(shapes)
        node_index = len(self._inbound)
        self._inbound.append(inputs)
        if isinstance(out_shape, list):
            return [KTensor(s, self, node_index, i) for i, s in enumerate(out_shape)]
        return KTensor(out_shape, self, node_index, 0)

    def get_weights(self):
        return [w.detach().numpy().astype(np.float32 if _FLOATX == "float32" else np.float64).copy()
                for w in self.trainable_weights]

    def set_weights(self, weights):
        if len(weights) != len(self.trainable_weights):
            raise ValueError(f'You called `set_weights(weights)` on layer "{self.name}" with a  weight list of length '
                             f"{len(weights)}, but the layer was expecting {len(self.trainable_weights)} weights.")
        for p, w in zip(self.trainable_weights, weights):
            if tuple(p.shape) != tuple(np.shape(w)):
                raise ValueError(f"Layer weight shape {tuple(p.shape)} not compatible with provided weight shape {np.shape(w)}")
            with torch.no_grad():
                p.copy_(torch.as_tensor(np.asarray(w), dtype=p.dtype))


class InputLayer(Layer):
    pass


def Input(shape=None, name=None, **unused):
    layer = InputLayer(name=name or "input_" + str(get_uid("input")))
    layer.built = True
    layer._inbound.append(None)
    return KTensor((None,) + tuple(shape), layer, 0, 0)


class Dense(Layer):                                              # layers/core.py Dense
    def __init__(self, units, activation=None, use_bias=True, **kwargs):
        super().__init__(**kwargs)
        self.units = int(units)
        self.activation = activations.get(activation)
        self.use_bias = use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight(name="kernel", shape=(input_shape[-1], self.units), initializer="glorot_uniform")
        self.bias = self.add_weight(name="bias", shape=(self.units,), initializer="zeros") if self.use_bias else None
        self.built = True

    def call(self, inputs):
        output = dot(inputs, self.kernel)
        if self.use_bias:
            output = bias_add(output, self.bias, data_format="channels_last")
        return self.activation(output)

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[:-1]) + (self.units,)


class Concatenate(Layer):                                        # layers/merge.py Concatenate
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, inputs):
        return concatenate(inputs, axis=self.axis)

    def compute_output_shape(self, input_shape):
        out = list(input_shape[0])
        out[self.axis] = sum(s[self.axis] for s in input_shape)
        return tuple(out)


def concatenate_layer(inputs, axis=-1, **kwargs):                # keras.layers.concatenate
    return Concatenate(axis=axis, **kwargs)(inputs)


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super().__init__(**kwargs)
        self.function = function

    def call(self, inputs):
        return self.function(inputs)


def add(inputs, **kwargs):
    raise NotImplementedError("keras.layers.add is imported by BS_brain.py but never called")


# ------------------------------------------------------------------ backend ops (backend/tensorflow_backend.py)
def ndim(x):
    return x.dim()


def dot(x, y):
    return torch.matmul(x, y)


def bias_add(x, bias, data_format=None):
    assert data_format in (None, "channels_last")
    return x + bias


def concatenate(tensors, axis=-1):
    return torch.cat(list(tensors), dim=axis)


def batch_dot(x, y, axes=None):
    """backend/tensorflow_backend.py batch_dot (2.2.4): pad the lower-rank operand with trailing 1-dims, one batched
    matmul with the adjoint flags derived from `axes`, squeeze the padding away."""
    if isinstance(axes, int):
        axes = (axes, axes)
    x_ndim, y_ndim = ndim(x), ndim(y)
    if axes is None:
        axes = [x_ndim - 1, y_ndim - 2]
    if x_ndim > y_ndim:
        diff = x_ndim - y_ndim
        y = y.reshape(tuple(y.shape) + (1,) * diff)
    elif y_ndim > x_ndim:
        diff = y_ndim - x_ndim
        x = x.reshape(tuple(x.shape) + (1,) * diff)
    else:
        diff = 0
    if ndim(x) == 2 and ndim(y) == 2:
        out = (x * y).sum(1) if axes[0] == axes[1] else (x.transpose(1, 0) * y).sum(1)
    else:
        adj_x = None if axes[0] == ndim(x) - 1 else True
        adj_y = True if axes[1] == ndim(y) - 1 else None
        out = torch.matmul(x.transpose(-1, -2) if adj_x else x, y.transpose(-1, -2) if adj_y else y)
    if diff:
        idx = x_ndim + y_ndim - 3 if x_ndim > y_ndim else x_ndim - 1
        for _ in range(diff):
            out = out.squeeze(idx)
    if ndim(out) == 1:
        out = out.unsqueeze(1)
    return out


# ------------------------------------------------------------------ optimizer (optimizers.py Adam, 2.2.4)
class Adam:
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0.0, amsgrad=False, **unused):
        assert decay == 0.0 and not amsgrad
        f = np.float32 if _FLOATX == "float32" else np.float64   # hyper-parameters are K.variable(floatx) in Keras
        self.lr, self.beta_1, self.beta_2 = float(f(lr)), float(f(beta_1)), float(f(beta_2))
        self.epsilon = _EPSILON if epsilon is None else epsilon
        self.iterations = 0
        self.ms = self.vs = None

    def apply(self, params, grads):
        """get_updates: t = iterations + 1; lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);
        m_t = b1 m + (1 - b1) g; v_t = b2 v + (1 - b2) g^2; p_t = p - lr_t m_t / (sqrt(v_t) + epsilon)."""
        if self.ms is None:
            self.ms = [torch.zeros_like(p) for p in params]
            self.vs = [torch.zeros_like(p) for p in params]
        self.iterations += 1
        dt = params[0].dtype
        t = torch.tensor(float(self.iterations), dtype=dt)
        b1, b2, lr = (torch.tensor(v, dtype=dt) for v in (self.beta_1, self.beta_2, self.lr))
        lr_t = lr * (torch.sqrt(1.0 - torch.pow(b2, t)) / (1.0 - torch.pow(b1, t)))
        with torch.no_grad():
            for p, g, m, v in zip(params, grads, self.ms, self.vs):
                m_t = b1 * m + (1.0 - b1) * g
                v_t = b2 * v + (1.0 - b2) * g * g
                p.sub_(lr_t * m_t / (torch.sqrt(v_t) + self.epsilon))
                m.copy_(m_t)
                v.copy_(v_t)


# ------------------------------------------------------------------ Model (engine/network.py + engine/training.py)
class History:
    def __init__(self):
        self.epoch, self.history = [], {}


class Model:
    def __init__(self, inputs, outputs, name=None):
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        self.name = name or "model_" + str(get_uid("model"))
        self.input_names = [t._history[0].name for t in self.inputs]
        self.output_names = [t._history[0].name for t in self.outputs]
        # layers in a deterministic topological order (inputs first, then by first use on the way to the outputs)
        self.layers, seen = [], set()

        def visit(t):
            layer, node, _ = t._history
            if (id(layer), node) in seen:
                return
            seen.add((id(layer), node))
            ins = layer._inbound[node]
            if ins is not None:
                for u in (ins if isinstance(ins, (list, tuple)) else [ins]):
                    visit(u)
            if layer not in self.layers:
                self.layers.append(layer)

        for t in self.inputs:
            visit(t)
        for t in self.outputs:
            visit(t)
        self.optimizer = self.loss = None

    # --- graph execution
    def _run(self, feed):
        cache = {}

        def value(t):
            layer, node, idx = t._history
            key = (id(layer), node)
            if key not in cache:
                ins = layer._inbound[node]
                if ins is None:
                    cache[key] = [feed[layer.name]]
                else:
                    args = [value(u) for u in ins] if isinstance(ins, (list, tuple)) else value(ins)
                    out = layer.call(args)
                    cache[key] = list(out) if isinstance(out, (list, tuple)) else [out]
            return cache[key][idx]

        return [value(t) for t in self.outputs]

    def _standardize(self, data, names):
        if isinstance(data, dict):
            missing = [n for n in names if n not in data]
            if missing:
                raise ValueError(f'No data provided for "{missing[0]}". Need data for each key in: {names}')
            arrays = [data[n] for n in names]
        else:
            arrays = list(data) if isinstance(data, (list, tuple)) else [data]
        return [torch.as_tensor(np.asarray(a), dtype=_tdtype()) for a in arrays]    # feed_dict casts to the placeholder dtype

    @property
    def trainable_weights(self):
        return [w for layer in self.layers for w in layer.trainable_weights]

    def get_layer(self, name=None, index=None):
        if index is not None:
            return self.layers[index]
        for layer in self.layers:
            if layer.name == name:
                return layer
        raise ValueError("No such layer: " + str(name))

    def get_weights(self):
        return [w for layer in self.layers for w in layer.get_weights()]

    def set_weights(self, weights):
        i = 0
        for layer in self.layers:
            n = len(layer.trainable_weights)
            layer.set_weights(weights[i:i + n])
            i += n
        assert i == len(weights)

    def save_weights(self, path):
        np.savez(path if path.endswith(".npz") else path + ".npz", *self.get_weights())

    def load_weights(self, path):
        z = np.load(path if path.endswith(".npz") else path + ".npz")
        self.set_weights([z[f"arr_{i}"] for i in range(len(z.files))])

    # --- inference / training
    def predict(self, x, batch_size=None, verbose=0):
        batch_size = batch_size or 32                             # training.py predict: default batch_size 32
        xs = self._standardize(x, self.input_names)
        n = xs[0].shape[0]
        outs = [[] for _ in self.outputs]
        with torch.no_grad():
            for s in range(0, n, batch_size):
                feed = {name: a[s:s + batch_size] for name, a in zip(self.input_names, xs)}
                for o, v in zip(outs, self._run(feed)):
                    o.append(v)
        res = [torch.cat(o, 0).numpy() for o in outs]
        return res if len(res) > 1 else res[0]

    def compile(self, optimizer, loss, **unused):
        self.optimizer, self.loss = optimizer, loss

    def _weighted_loss(self, y_true, y_pred):
        """engine/training_utils.py weighted_masked_objective with the all-ones sample weights `fit` feeds by default."""
        score = self.loss(y_true, y_pred)
        weights = torch.ones(y_true.shape[0], dtype=y_pred.dtype)
        if score.dim() == 0:
            score = score * weights                               # a scalar loss broadcasts against the (batch,) weights
        else:
            score = score.reshape(score.shape[0], -1).mean(1) * weights if score.dim() > 1 else score * weights
        score = score / (weights != 0).to(score.dtype).mean()
        return score.mean()

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, shuffle=True, **unused):
        batch_size = batch_size or 32
        xs = self._standardize(x, self.input_names)
        ys = self._standardize(y, self.output_names)
        n = xs[0].shape[0]
        hist = History()
        names = ["loss"] + ([o + "_loss" for o in self.output_names] if len(self.outputs) > 1 else [])
        params = self.trainable_weights
        for epoch in range(epochs):
            index = np.arange(n)
            if shuffle:
                np.random.shuffle(index)                          # training_arrays.fit_loop: the GLOBAL numpy generator
            totals = np.zeros(len(names))
            for s in range(0, n, batch_size):
                ids = torch.as_tensor(index[s:s + batch_size])
                feed = {name: a[ids] for name, a in zip(self.input_names, xs)}
                preds = self._run(feed)
                per_out = [self._weighted_loss(t[ids], p) for t, p in zip(ys, preds)]
                total = sum(per_out)
                grads = torch.autograd.grad(total, params, allow_unused=True)
                grads = [torch.zeros_like(p) if g is None else g for p, g in zip(params, grads)]
                self.optimizer.apply(params, grads)
                vals = [float(total.detach())] + ([float(v.detach()) for v in per_out] if len(self.outputs) > 1 else [])
                totals += np.array(vals) * len(ids)               # callbacks.BaseLogger: batch-size-weighted epoch mean
            hist.epoch.append(epoch)
            for k, v in zip(names, totals / n):
                hist.history.setdefault(k, []).append(float(v))
        return hist

    def train_on_batch(self, x, y):
        return self.fit(x, y, batch_size=len(next(iter(y.values())) if isinstance(y, dict) else y[0]), shuffle=False).history["loss"][0]

"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): a numpy fp64 stand-in for the handful of `tensorflow` /
`tensorflow.keras` names that the reference's LAYER files use, so that those files can be imported and EXECUTED
unmodified where TensorFlow is not installable (tools/gen_reference_layer_golden.py, run in the container that has
/root/reference).  Their outputs on seeded inputs / weights are committed as tests/golden/ref_layers.npz and
tests/test_oracle_reference_pin.py checks oracle/oracle_np.py + oracle/oracle_models.py against them: the oracle is
then pinned to the reference's own code path (its order of splits, concats, masks, residuals), with only the meaning of
each individual TF op restated here.  Each op follows its documented TensorFlow semantics:

  tf.concat / split (n equal parts, or a list of sizes) / transpose(perm) / reshape / tile / squeeze / expand_dims / slice / where / equal /
  zeros_like / ones_like / square / identity / clip_by_value / shape / newaxis / tensordot / matmul (batched, transpose_b)
  tf.math.reduce_sum(axis, keepdims); tf.nn.softmax (last axis) / relu / sigmoid / bias_add; tf.sequence_mask
  tf.keras.layers.Layer (build on first call, add_weight), Dense, Dropout (inference: identity), Activation, Flatten,
  Concatenate; tf.keras.initializers.* / regularizers.* (accepted, ignored: weights are drawn from the shim's seeded RNG,
  biases included, so that no term is trivially zero)

For the composed graph (staytime/VideoDnn.py::create_moe_sub_model) the Keras FUNCTIONAL code is run eagerly:
`tn.layers.Input(name=...)` hands back the concrete seeded array registered under that name (FEEDS), so "building" the
model computes it; `tn.model.Model` just keeps the outputs.  Extra names for that file: tf.multiply / stack / constant /
stop_gradient (identity) / float32 (fp64 here), tf.keras.layers.multiply / ReLU / Lambda.

Nothing in the product imports this module."""
from __future__ import annotations

import sys
import types

import numpy as np

_RNG = np.random.default_rng(0)


def seed(s):
    global _RNG
    _RNG = np.random.default_rng(s)


class _Shape(tuple):
    def as_list(self):
        return list(self)


class Tensor(np.ndarray):
    """ndarray with the two TensorShape accessors the reference calls."""

    def get_shape(self):
        return _Shape(self.shape)


def T(a, dtype=None):
    a = np.asarray(a, dtype=dtype)
    if a.dtype.kind == "f":
        a = a.astype(np.float64)
    return a.view(Tensor)


def _axis(axis):
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


# ---------------------------------------------------------------------------------------------- ops
def concat(values, axis):
    return T(np.concatenate([np.asarray(v) for v in values], axis=axis))


def split(value, num_or_size_splits, axis=0):
    if isinstance(num_or_size_splits, (list, tuple)):          # tf: a list holds the SIZES of the parts
        num_or_size_splits = np.cumsum([int(n) for n in num_or_size_splits])[:-1]
    return [T(p) for p in np.split(np.asarray(value), num_or_size_splits, axis=axis)]


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = np.asarray(a), np.asarray(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return T(np.matmul(a, b))


def tensordot(a, b, axes):
    if isinstance(axes, (tuple, list)):
        ax = ([axes[0]] if np.isscalar(axes[0]) else list(axes[0]), [axes[1]] if np.isscalar(axes[1]) else list(axes[1]))
    else:
        ax = axes
    return T(np.tensordot(np.asarray(a), np.asarray(b), axes=ax))


def transpose(a, perm=None):
    return T(np.transpose(np.asarray(a), perm))


def reshape(t, shape):
    return T(np.reshape(np.asarray(t), tuple(int(s) for s in np.asarray(shape).reshape(-1))))


def tile(t, multiples):
    return T(np.tile(np.asarray(t), tuple(int(m) for m in multiples)))


def squeeze(t, axis=None):
    return T(np.squeeze(np.asarray(t), axis=None if axis is None else _axis(axis)))


def expand_dims(t, axis):
    ax = _axis(axis)
    return T(np.expand_dims(np.asarray(t), ax))


def slice_(t, begin, size):
    t = np.asarray(t)
    idx = tuple(slice(int(b), t.shape[i] if int(s) == -1 else int(b) + int(s)) for i, (b, s) in enumerate(zip(begin, size)))
    return T(t[idx])


def where(cond, x, y, name=None):
    return T(np.where(np.asarray(cond).astype(bool), np.asarray(x), np.asarray(y)))


def shape(t):
    return np.asarray(np.asarray(t).shape, dtype=np.int64)


def sequence_mask(lengths, maxlen=None):
    lengths = np.asarray(lengths)
    maxlen = int(lengths.max()) if maxlen is None else int(maxlen)
    return T(np.arange(maxlen)[None, :] < lengths[..., None])


def reduce_sum(t, axis=None, keepdims=False):
    return T(np.sum(np.asarray(t), axis=None if axis is None else _axis(axis), keepdims=keepdims))


def softmax(t, axis=-1):
    t = np.asarray(t)
    e = np.exp(t - t.max(axis=axis, keepdims=True))
    return T(e / e.sum(axis=axis, keepdims=True))


def relu(t):
    return T(np.maximum(np.asarray(t), 0.0))


def sigmoid(t):
    return T(1.0 / (1.0 + np.exp(-np.asarray(t))))


_ACT = {None: lambda t: T(t), "linear": lambda t: T(t), "relu": relu, "sigmoid": sigmoid, "softmax": softmax}


def _activation(a):
    return a if callable(a) else _ACT[a]


# ---------------------------------------------------------------------------------------------- keras
LAYERS = []        # every layer the executed reference code created, in creation order (weights are read back by name)
FEEDS = {}         # name -> array handed out by tn.layers.Input(name=...)
PARSED = {}        # what tf.io.parse_example returns (staytime/parse.py)
WEIGHT_LOG = []    # every weight array in creation order: with the seed, a run's weights can be re-drawn (replay_weights)


def draw_weight(rng, shape):
    """The shim's one initialiser: matrices ~ N(0, 1/fan_in) so that deep stacks keep O(1) activations (no saturated
    softmax hiding an error), everything else (biases, gamma - 1, beta, vectors) ~ N(0, 0.3^2)."""
    shape = tuple(int(s) for s in shape)
    scale = 1.0 / np.sqrt(shape[0]) if len(shape) == 2 and shape[0] > 1 else 0.3
    return rng.standard_normal(shape) * scale


def replay_weights(seed_, manifest):
    """Re-draws the weights of a run from its seed and its manifest [[key, shape, offset], ...] (creation order; key ""
    = a weight the oracle has no name for): the fixtures store the manifest instead of megabytes of random numbers."""
    rng = np.random.default_rng(int(seed_))
    P = {}
    for key, shape, offset in manifest:
        w = draw_weight(rng, shape) + float(offset)
        if key:
            P[key] = w
    return P


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name
        self.built = False
        self._w = {}
        LAYERS.append(self)

    def build(self, input_shape):
        self.built = True

    def add_weight(self, name=None, shape=None, initializer=None, regularizer=None, trainable=True, **kw):
        w = T(draw_weight(_RNG, shape))
        self._w[name] = w
        WEIGHT_LOG.append(w)
        return w

    def __call__(self, *args, **kwargs):
        if not self.built:
            first = args[0]
            self.build(_Shape(np.asarray(first[0] if isinstance(first, (list, tuple)) else first).shape))
            self.built = True
        return self.call(*args, **kwargs)

    def get_config(self):
        return {}


class Dense(Layer):
    def __init__(self, units, activation=None, name=None, **kwargs):
        super().__init__(name=name)
        self.units, self.activation = int(units), _activation(activation)

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (int(input_shape[-1]), self.units))
        self.bias = self.add_weight("bias", (self.units,))

    def call(self, x):
        return self.activation(T(np.matmul(np.asarray(x), self.kernel) + self.bias))


class Dropout(Layer):
    def __init__(self, rate, seed=None, **kwargs):
        super().__init__()
        self.rate = rate

    def call(self, x, training=None):
        return T(x)            # inference: identity


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__()
        self.fn = _activation(activation)

    def call(self, x):
        return self.fn(x)


class Flatten(Layer):
    def call(self, x):
        x = np.asarray(x)
        return T(x.reshape(x.shape[0], -1))


class Concatenate(Layer):
    def __init__(self, axis=-1, name=None, **kwargs):
        super().__init__(name=name)
        self.axis = axis

    def call(self, xs):
        return concat(xs, self.axis)


class ReLU(Layer):
    def call(self, x):
        return relu(x)


class Lambda(Layer):
    def __init__(self, function, name=None, **kwargs):
        super().__init__(name=name)
        self.function = function

    def build(self, input_shape):
        self.built = True

    def __call__(self, x):
        return self.function(x)


class Model:
    """tn.model.Model / tf.keras.Model of an eagerly executed functional graph: keeps the outputs; calling it (the
    reference applies a sub-model to the tensors its Inputs were registered with) hands them back."""

    def __init__(self, inputs=None, outputs=None, name=None, **kwargs):
        self.inputs, self.outputs, self.output, self.name = inputs, outputs, outputs, name

    def __call__(self, *args, **kwargs):
        return self.outputs


def identity(t, name=None):
    """tf.identity.  A NAMED result is also registered as FEEDS["shallow_<name>"]: rough_rank/model.py feeds the towers'
    named outputs to create_shallow_tower's Inputs of exactly that name (:72-73, :157-163)."""
    t = T(t)
    if name is not None:
        t.name = name + ":0"
        FEEDS.setdefault("shallow_" + name, t)
    return t


def cast(t, dtype):
    a = np.asarray(t)
    return T(a.astype(bool)) if dtype in ("bool", bool) else T(a.astype(np.float64))


class MeanSquaredError:
    """tf.keras.losses.MeanSquaredError(reduction=NONE): mean over the last axis."""

    def __init__(self, reduction=None, **kw):
        pass

    def __call__(self, y_true, y_pred):
        return T(np.mean((np.asarray(y_pred) - np.asarray(y_true)) ** 2, axis=-1))


def Input(name=None, **kwargs):
    v = FEEDS.get(name)
    return T(np.zeros((0,))) if v is None else T(v)       # id inputs feed nothing in the dense graphs: placeholder


class LayerNormalization(Layer):
    """The reference imports `.layer_normalization.LayerNormalization`, a file that is NOT in the reference tree
    (SURVEY.md §8): normalisation over the last axis with learned gamma / beta, variance epsilon `eps`."""

    def __init__(self, eps=1e-3, **kwargs):
        super().__init__()
        self.eps = eps

    def build(self, input_shape):
        self.gamma = self.add_weight("gamma", (int(input_shape[-1]),)) + 1.0
        self.beta = self.add_weight("beta", (int(input_shape[-1]),))

    def call(self, x):
        x = np.asarray(x)
        mean = x.mean(-1, keepdims=True)
        var = ((x - mean) ** 2).mean(-1, keepdims=True)
        return T((x - mean) / np.sqrt(var + self.eps) * self.gamma + self.beta)


def install():
    """Registers the stand-in as `tensorflow` (+ the submodules the reference imports) in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.concat, tf.split, tf.matmul, tf.tensordot, tf.transpose, tf.reshape = concat, split, matmul, tensordot, transpose, reshape
    tf.tile, tf.squeeze, tf.expand_dims, tf.slice, tf.where, tf.shape = tile, squeeze, expand_dims, slice_, where, shape
    tf.sequence_mask = sequence_mask
    tf.equal = lambda a, b: T(np.asarray(a) == np.asarray(b))
    tf.zeros_like = lambda t: T(np.zeros_like(np.asarray(t)))
    tf.ones_like = lambda t: T(np.ones_like(np.asarray(t)))
    tf.square = lambda t: T(np.square(np.asarray(t)))
    tf.identity = identity
    tf.sigmoid = sigmoid
    tf.add = lambda a, b: T(np.asarray(a) + np.asarray(b))
    tf.greater = lambda a, b: T(np.asarray(a) > np.asarray(b))
    tf.clip_by_value = lambda t, lo, hi: T(np.clip(np.asarray(t), lo, hi))
    tf.newaxis = None
    tf.multiply = lambda a, b, name=None: T(np.asarray(a) * np.asarray(b))
    tf.stack = lambda xs, axis=0: T(np.stack([np.asarray(x) for x in xs], axis=axis))
    tf.constant = lambda v, dtype=None: T(np.asarray(v, dtype=np.float64))
    tf.stop_gradient = lambda t: T(t)
    tf.float32 = np.float64                 # everything runs in fp64 here
    reduce_mean = lambda t, axis=None, keepdims=False: T(np.mean(np.asarray(t), axis=None if axis is None else _axis(axis),
                                                                keepdims=keepdims))
    tf.math = types.SimpleNamespace(reduce_sum=reduce_sum, reduce_mean=reduce_mean, log=lambda t: T(np.log(np.asarray(t))))
    tf.reduce_sum, tf.reduce_mean = reduce_sum, reduce_mean
    tf.cast = cast                                                       # labels -> float (fp64 here), masks -> bool
    tf.abs = lambda t: T(np.abs(np.asarray(t)))
    tf.summary = types.SimpleNamespace(scalar=lambda *a, **k: None)
    # staytime/parse.py::parse_input_func: the parsed example is handed in ready (PARSED), the label arithmetic runs here
    import re as _re
    tf.io = types.SimpleNamespace(FixedLenFeature=lambda *a, **k: None, VarLenFeature=lambda *a, **k: None,
                                  parse_example=lambda proto, desc: dict(PARSED))
    tf.string, tf.int64 = "string", "int64"
    tf.divide = lambda a, b: T(np.asarray(a, np.float64) / np.asarray(b, np.float64))
    tf.repeat = lambda t, n, axis=None: T(np.repeat(np.asarray(t), int(n), axis=axis))
    tf.math.subtract = lambda a, b: T(np.asarray(a) - np.asarray(b))
    tf.math.square = lambda t: T(np.square(np.asarray(t)))
    tf.math.abs = lambda t: T(np.abs(np.asarray(t)))
    tf.math.exp = lambda t: T(np.exp(np.asarray(t)))
    tf.strings = types.SimpleNamespace(regex_full_match=lambda t, pat: T(np.vectorize(
        lambda v: _re.fullmatch(pat, v.decode() if isinstance(v, bytes) else str(v)) is not None)(np.asarray(t))))
    tf.nn = types.SimpleNamespace(softmax=softmax, relu=relu, sigmoid=sigmoid,
                                  bias_add=lambda x, b: T(np.asarray(x) + np.asarray(b)))
    layers = types.ModuleType("tensorflow.keras.layers")
    for cls in (Layer, Dense, Dropout, Activation, Flatten, Concatenate, ReLU, Lambda):
        setattr(layers, cls.__name__, cls)
    layers.multiply = lambda xs: T(np.asarray(xs[0]) * np.asarray(xs[1]))
    layers.Input = Input
    keras = types.ModuleType("tensorflow.keras")
    keras.layers = layers
    keras.initializers = types.SimpleNamespace(GlorotNormal=lambda seed=None: None, Zeros=lambda: None,
                                               TruncatedNormal=lambda **kw: None)
    keras.regularizers = types.SimpleNamespace(L2=lambda l2=0: None, l2=lambda l2=0: None, L1L2=lambda **kw: None)
    for unused in ("Embedding", "BatchNormalization"):      # imported by rank/multi_head/multidnn.py, never called
        setattr(layers, unused, type(unused, (Layer,), {}))
    tf.keras = keras
    python = types.ModuleType("tensorflow.python")
    pkeras = types.ModuleType("tensorflow.python.keras")
    backend = types.ModuleType("tensorflow.python.keras.backend")
    backend.ndim = lambda t: np.asarray(t).ndim
    backend.epsilon = lambda: 1e-7                                       # tf.keras.backend.epsilon() default
    backend.clip = lambda t, lo, hi: T(np.clip(np.asarray(t), lo, hi))
    pkeras.backend = backend
    pkeras.layers = layers
    callbacks = types.ModuleType("tensorflow.python.keras.callbacks")
    callbacks.Callback = type("Callback", (), {})
    pkeras.callbacks = callbacks
    python.keras = pkeras
    tf.python = python
    keras.Model = Model
    keras.losses = types.SimpleNamespace(MeanSquaredError=MeanSquaredError, Reduction=types.SimpleNamespace(NONE="none"))
    # tensornet (the reference's parameter-server framework): only what the dense sub-graph builders touch
    tn = types.ModuleType("tensornet")
    tn.layers = types.SimpleNamespace(Input=Input)
    tn.model = types.SimpleNamespace(Model=Model)
    for name, mod in (("tensorflow", tf), ("tensorflow.keras", keras), ("tensorflow.keras.layers", layers),
                      ("tensorflow.python", python), ("tensorflow.python.keras", pkeras),
                      ("tensorflow.python.keras.backend", backend), ("tensorflow.python.keras.layers", layers),
                      ("tensorflow.python.keras.callbacks", callbacks), ("tensornet", tn)):
        sys.modules[name] = mod
    return tf

"""The numpy stand-in for TensorFlow (oracle/tf_numpy_shim.py) that the reference's files are executed under when the
golden vectors are generated: every op it restates is checked here against an INDEPENDENT implementation of the same
documented semantics (torch on CPU, or plain index arithmetic), so that a pin obtained through the shim cannot be a
shared misreading of an op."""
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def tf():
    from oracle import tf_numpy_shim as shim
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "tensorflow" or k.startswith("tensorflow.") or k == "tensornet"}
    mod = shim.install()
    yield mod
    for k in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.") or k == "tensornet"]:
        del sys.modules[k]
    sys.modules.update({k: v for k, v in saved.items() if v is not None})


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_shape_ops(tf):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 6, 8))
    parts = tf.split(x, 2, axis=2)                                  # n equal parts
    assert all(np.array_equal(p, q.numpy()) for p, q in zip(parts, torch.chunk(t(x), 2, dim=2)))
    sizes = tf.split(x, [1, 3, 4], axis=2)                          # a list = the SIZES of the parts
    assert [p.shape[2] for p in sizes] == [1, 3, 4]
    assert all(np.array_equal(p, q.numpy()) for p, q in zip(sizes, torch.split(t(x), [1, 3, 4], dim=2)))
    assert np.array_equal(tf.concat(parts, axis=0), torch.cat(list(torch.chunk(t(x), 2, dim=2)), 0).numpy())
    assert np.array_equal(tf.transpose(x, [0, 2, 1]), t(x).permute(0, 2, 1).numpy())
    assert np.array_equal(tf.tile(x[:, :1], [1, 3, 2]), t(x[:, :1]).repeat(1, 3, 2).numpy())
    assert np.array_equal(tf.reshape(x, [-1, 1, 8]), x.reshape(-1, 1, 8))
    assert tf.expand_dims(x, axis=[1]).shape == (4, 1, 6, 8) and tf.squeeze(x[:, :1], [1]).shape == (4, 8)
    assert np.array_equal(tf.slice(x, [0, 1, 2], [2, 3, -1]), x[0:2, 1:4, 2:])
    assert np.array_equal(tf.stack([x, x + 1], axis=1), torch.stack([t(x), t(x) + 1], 1).numpy())
    assert np.array_equal(tf.repeat(x[:, :1, 0], 5, axis=1), t(x[:, :1, 0]).repeat_interleave(5, dim=1).numpy())
    assert list(tf.shape(x)) == [4, 6, 8] and x.view(type(tf.identity(x))).get_shape().as_list() == [4, 6, 8]


def test_math_ops(tf):
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((3, 5, 7)), rng.standard_normal((3, 7, 2))
    assert np.allclose(tf.matmul(a, b), torch.matmul(t(a), t(b)).numpy(), atol=1e-13)
    assert np.allclose(tf.matmul(a, np.swapaxes(b, 1, 2), transpose_b=True), torch.matmul(t(a), t(b)).numpy(), atol=1e-13)
    k = rng.standard_normal((5, 1))
    assert np.allclose(tf.tensordot(a, k, axes=(1, 0)), torch.tensordot(t(a), t(k), dims=([1], [0])).numpy(), atol=1e-13)
    assert np.allclose(tf.tensordot(a, b[0], axes=(-1, 0)), torch.tensordot(t(a), t(b[0]), dims=([2], [0])).numpy(), atol=1e-13)
    assert np.allclose(tf.nn.softmax(a), torch.softmax(t(a), -1).numpy(), atol=1e-15)
    assert np.allclose(tf.nn.sigmoid(a), torch.sigmoid(t(a)).numpy(), atol=1e-15)
    assert np.array_equal(tf.nn.relu(a), torch.relu(t(a)).numpy())
    assert np.allclose(tf.math.reduce_sum(a, axis=1, keepdims=True), t(a).sum(1, keepdim=True).numpy(), atol=1e-13)
    assert np.allclose(tf.reduce_sum([a, a * 2], axis=0), 3 * a, atol=1e-13)           # a python list of tensors
    assert np.allclose(tf.math.reduce_mean(a, axis=0), t(a).mean(0).numpy(), atol=1e-13)
    assert np.array_equal(tf.where(a > 0, a, tf.zeros_like(a)), torch.where(t(a) > 0, t(a), torch.zeros_like(t(a))).numpy())
    assert np.array_equal(tf.clip_by_value(a, -0.5, 0.5), t(a).clamp(-0.5, 0.5).numpy())
    m = tf.sequence_mask(np.array([0, 2, 4]))
    assert m.shape == (3, 4) and m.tolist() == [[False] * 4, [True, True, False, False], [True] * 4]
    assert tf.cast(np.array([0.0, 1.0]) == 1, dtype="bool").dtype == bool
    assert tf.strings.regex_full_match(np.array(["ab_video_homepage_landing_c", "label"]), ".*video_homepage_landing.*").tolist() == [True, False]


def test_keras_layers(tf):
    from oracle import tf_numpy_shim as shim
    shim.seed(3)
    rng = np.random.default_rng(2)
    x = rng.standard_normal((5, 9))
    d = tf.keras.layers.Dense(4, activation="relu", name="d")
    y = d(x)
    assert np.allclose(y, torch.relu(t(x) @ t(d.kernel) + t(d.bias)).numpy(), atol=1e-13) and d.kernel.shape == (9, 4)
    assert np.array_equal(d(x), y)                                   # built once: the same weights on every call
    ln = shim.LayerNormalization(eps=1e-3)
    z = ln(x)
    ref = torch.nn.functional.layer_norm(t(x), (9,), t(np.asarray(ln.gamma)), t(np.asarray(ln.beta)), 1e-3)
    assert np.allclose(z, ref.numpy(), atol=1e-13)
    assert np.array_equal(tf.keras.layers.Concatenate(axis=1)([x, x]), np.concatenate([x, x], 1))
    assert np.array_equal(tf.keras.layers.Flatten()(x.reshape(5, 3, 3)), x)
    assert np.array_equal(tf.keras.layers.multiply([x, x]), x * x)
    mse = tf.keras.losses.MeanSquaredError(reduction=tf.keras.losses.Reduction.NONE)
    assert np.allclose(mse(x, x + 1.0), np.ones(5))
    # replay_weights re-draws exactly what the layers drew
    shim.seed(7)
    del shim.WEIGHT_LOG[:]
    e = tf.keras.layers.Dense(3)
    e(x)
    P = shim.replay_weights(7, [["k", [9, 3], 0.0], ["b", [3], 0.0]])
    assert np.array_equal(P["k"], e.kernel) and np.array_equal(P["b"], e.bias)

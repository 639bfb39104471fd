"""The TensorFlow stand-in of tests/golden/tf1_shim against independent implementations.

The golden generator executes the reference's scripts over this stand-in, so the fixtures are only as good as its
primitives.  Each one is held here to (a) PyTorch's own operator where PyTorch implements the same definition
(convolution, pooling, align_corners bilinear, batch-stat normalisation, softmax cross-entropy) with TensorFlow's
documented padding rule applied by hand, and (b) the formula TensorFlow's documentation gives where PyTorch has no
counterpart (weighted_cross_entropy_with_logits, 'SAME' padding split, legacy resize index rules), plus the reference's
vendored slim golden vectors for 'SAME' convolutions.  float64, exact to round-off.
"""
import importlib.util
import json
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def tf():
    """The stand-in under a private module name (it must not become `tensorflow` for the rest of the test session)."""
    root = os.path.join(HERE, "golden", "tf1_shim", "tensorflow")
    spec = importlib.util.spec_from_file_location("basi_tf1_shim", os.path.join(root, "__init__.py"),
                                                  submodule_search_locations=[root])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["basi_tf1_shim"] = mod
    spec.loader.exec_module(mod)
    yield mod
    for k in [k for k in sys.modules if k == "basi_tf1_shim" or k.startswith("basi_tf1_shim.")]:
        del sys.modules[k]


def T(tf, a):
    return tf.Tensor(torch.as_tensor(np.asarray(a, dtype=np.float64)))


def nchw(t):
    return t.permute(0, 3, 1, 2)


def same_pad(n, k, s, d=1):
    """TensorFlow's documented rule: out = ceil(n / s); total = max((out-1)*s + (k-1)*d+1 - n, 0); before = total // 2."""
    out = -(-n // s)
    total = max((out - 1) * s + (k - 1) * d + 1 - n, 0)
    return total // 2, total - total // 2


@pytest.mark.parametrize("case", [(3, 1, 1, "SAME", 9, 11), (3, 2, 1, "SAME", 10, 10), (3, 2, 1, "SAME", 9, 12),
                                  (1, 2, 1, "VALID", 9, 10), (5, 5, 1, "VALID", 5, 5), (3, 1, 2, "VALID", 12, 13),
                                  (3, 1, 4, "SAME", 11, 9), (7, 1, 1, "VALID", 7, 8)])
def test_convolution_equals_pytorch_with_tf_padding(tf, case):
    k, s, d, padding, h, w = case
    rng = np.random.RandomState(0)
    x, wt = rng.randn(2, h, w, 5), rng.randn(k, k, 5, 4)
    if d == 1:
        y = tf.nn.conv2d(T(tf, x), T(tf, wt), [1, s, s, 1], padding=padding).t
    else:
        y = tf.nn.atrous_conv2d(T(tf, x), T(tf, wt), d, padding=padding).t
    xt = nchw(torch.as_tensor(x))
    if padding == "SAME":
        (pt, pb), (pl, pr) = same_pad(h, k, s, d), same_pad(w, k, s, d)
        xt = F.pad(xt, (pl, pr, pt, pb))
    ref = F.conv2d(xt, torch.as_tensor(wt).permute(3, 2, 0, 1), stride=s, dilation=d).permute(0, 2, 3, 1)
    assert y.shape == ref.shape and float((y - ref).abs().max()) < 1e-12


def test_same_convolutions_reproduce_the_reference_vendored_slim_vectors(tf):
    """The numeric golden vectors of the reference's own vendored TF-slim tests (slim/nets/resnet_v1_test.py:72-153,
    extracted to tests/golden/slim_reference_tests.json): the 3x3 mesh filter on the n x n mesh image -- 'SAME' at stride
    1, its subsample, explicit padding + 'VALID' at stride 2, 'SAME' at stride 2 -- on an even and an odd input.  These
    are real TensorFlow outputs: the stand-in's convolution / padding rule is pinned to them, not only to PyTorch."""
    with open(os.path.join(HERE, "golden", "slim_reference_tests.json")) as f:
        g = json.load(f)["resnet_utils"]

    def mesh(n):
        return (np.arange(n).reshape(n, 1) + np.arange(n).reshape(1, n)).astype(np.float64)

    w = T(tf, mesh(3).reshape(3, 3, 1, 1))
    for case in ("conv2d_same_even", "conv2d_same_odd"):
        n = g[case]["n"]
        x = T(tf, mesh(n).reshape(1, n, n, 1))
        y1 = tf.nn.conv2d(x, w, [1, 1, 1, 1], padding="SAME")
        assert np.array_equal(y1.t.numpy()[0, :, :, 0], np.array(g[case]["y1_expected"], dtype=np.float64))
        assert np.array_equal(y1.t.numpy()[0, ::2, ::2, 0], np.array(g[case]["y2_expected"], dtype=np.float64))
        y3 = tf.nn.conv2d(tf.pad(x, [[0, 0], [1, 1], [1, 1], [0, 0]]), w, [1, 2, 2, 1], padding="VALID")
        assert np.array_equal(y3.t.numpy()[0, :, :, 0], np.array(g[case]["y3_expected"], dtype=np.float64))
        y4 = tf.nn.conv2d(x, w, [1, 2, 2, 1], padding="SAME")
        assert np.array_equal(y4.t.numpy()[0, :, :, 0], np.array(g[case]["y4_expected"], dtype=np.float64))


@pytest.mark.parametrize("case", [("max", 3, 2, "SAME", 10, 11), ("max", 2, 2, "VALID", 10, 12), ("avg", 3, 3, "VALID", 10, 10),
                                  ("avg", 2, 2, "VALID", 10, 7), ("avg", 10, 10, "VALID", 10, 10)])
def test_pooling_equals_pytorch(tf, case):
    kind, k, s, padding, h, w = case
    x = np.random.RandomState(1).randn(2, h, w, 3)
    fn = tf.nn.max_pool if kind == "max" else tf.nn.avg_pool
    y = fn(T(tf, x), [1, k, k, 1], [1, s, s, 1], padding).t
    xt = nchw(torch.as_tensor(x))
    if padding == "SAME":
        (pt, pb), (pl, pr) = same_pad(h, k, s), same_pad(w, k, s)
        xt = F.pad(xt, (pl, pr, pt, pb), value=-math.inf)
    ref = (F.max_pool2d(xt, k, s) if kind == "max" else F.avg_pool2d(xt, k, s)).permute(0, 2, 3, 1)
    assert y.shape == ref.shape and float((y - ref).abs().max()) < 1e-13


@pytest.mark.parametrize("size", [(1, 1, 10, 10), (2, 2, 10, 10), (3, 3, 10, 10), (5, 7, 12, 9), (10, 10, 10, 10)])
def test_align_corners_bilinear_equals_pytorch(tf, size):
    hi, wi, ho, wo = size
    x = np.random.RandomState(2).randn(2, hi, wi, 3)
    y = tf.image.resize_bilinear(T(tf, x), [ho, wo], align_corners=True).t
    ref = F.interpolate(nchw(torch.as_tensor(x)), size=(ho, wo), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    assert float((y - ref).abs().max()) < 1e-12


def test_legacy_resizes_follow_the_documented_index_rules(tf):
    """TF1 default (align_corners=False, no half-pixel centres): src = dst * in / out -- bilinear interpolates between
    floor(src) and floor(src) + 1 (clamped), nearest takes floor(src) computed in float32 like TF's kernel."""
    x = np.random.RandomState(3).randn(1, 5, 4, 2)
    y = tf.image.resize_bilinear(T(tf, x), [12, 9]).t.numpy()
    for oy in range(12):
        for ox in range(9):
            sy, sx = oy * 5 / 12.0, ox * 4 / 9.0
            y0, x0 = int(sy), int(sx)
            y1, x1 = min(y0 + 1, 4), min(x0 + 1, 3)
            fy, fx = sy - y0, sx - x0
            ref = ((1 - fy) * ((1 - fx) * x[0, y0, x0] + fx * x[0, y0, x1])
                   + fy * ((1 - fx) * x[0, y1, x0] + fx * x[0, y1, x1]))
            assert np.abs(y[0, oy, ox] - ref).max() < 1e-12
    lab = np.arange(224, dtype=np.float64).reshape(1, 224, 1, 1).repeat(2, axis=2)
    rows = tf.image.resize_nearest_neighbor(T(tf, lab), [110, 2]).t.numpy()[0, :, 0, 0]
    want = [min(int(np.floor(np.float32(o) * (np.float32(224) / np.float32(110)))), 223) for o in range(110)]
    assert rows.tolist() == want and rows[55] == 112          # float64 arithmetic would give 111


def test_batch_normalization_equals_pytorch_batch_stat_mode(tf):
    rng = np.random.RandomState(4)
    x = rng.randn(3, 4, 5, 6)
    vals = {"gamma": rng.uniform(0.5, 1.5, 6), "beta": rng.uniform(-1, 1, 6), "moving_mean": np.zeros(6),
            "moving_variance": np.ones(6)}
    tf.shim_reset(lambda n, s, k: vals[k], [])
    with tf.variable_scope("bn"):
        y = tf.layers.batch_normalization(T(tf, x), momentum=0.95, epsilon=1e-5, training=True, name="bn").t
    ref = F.batch_norm(nchw(torch.as_tensor(x)), None, None, torch.as_tensor(vals["gamma"]), torch.as_tensor(vals["beta"]),
                       training=True, eps=1e-5).permute(0, 2, 3, 1)
    assert float((y.detach() - ref).abs().max()) < 1e-12
    names = list(tf.shim_state().variables)
    assert names == ["bn/bn/gamma", "bn/bn/beta", "bn/bn/moving_mean", "bn/bn/moving_variance"]     # TF's doubled scope


def test_losses_follow_the_documented_formulas(tf):
    rng = np.random.RandomState(5)
    x, z = rng.randn(50) * 4, (rng.rand(50) > 0.6).astype(np.float64)
    for q in (1.0, 3.0, 5.0):
        got = tf.nn.weighted_cross_entropy_with_logits(targets=T(tf, z), logits=T(tf, x), pos_weight=q).t.numpy()
        # tf docs: (1 - z) * x + l * (log(1 + exp(-abs(x))) + max(-x, 0)),  l = 1 + (q - 1) * z
        ll = 1 + (q - 1) * z
        want = (1 - z) * x + ll * (np.log1p(np.exp(-np.abs(x))) + np.maximum(-x, 0))
        assert np.abs(got - want).max() < 1e-12
    logits, lab = rng.randn(7, 5), rng.randint(0, 5, size=7)
    got = tf.nn.sparse_softmax_cross_entropy_with_logits(labels=tf.Tensor(torch.as_tensor(lab)), logits=T(tf, logits)).t
    want = F.cross_entropy(torch.as_tensor(logits), torch.as_tensor(lab), reduction="none")
    assert float((got - want).abs().max()) < 1e-12


def test_minimize_is_plain_sgd_on_the_selected_variables(tf):
    vals = {"a/w": np.array([1.0, -2.0]), "b/w": np.array([0.5])}
    tf.shim_reset(lambda n, s, k: vals[n], [])
    with tf.variable_scope("a"):
        a = tf.get_variable("w", [2])
    with tf.variable_scope("b"):
        b = tf.get_variable("w", [1])
    loss = tf.reduce_mean(a * a) + 3.0 * tf.reduce_mean(b)
    op = tf.train.GradientDescentOptimizer(tf.constant(0.1)).minimize(loss)
    assert op.var_names == ["a/w", "b/w"]
    assert np.allclose(op.grads["a/w"], [1.0, -2.0]) and np.allclose(op.grads["b/w"], [3.0])
    assert np.allclose(op.new_values["a/w"], [0.9, -1.8]) and np.allclose(op.new_values["b/w"], [0.2])
    only_b = tf.train.GradientDescentOptimizer(0.1).minimize(loss, var_list=[v for v in tf.trainable_variables()
                                                                             if "b/" in v.name])
    assert only_b.var_names == ["b/w"]

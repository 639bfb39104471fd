"""Shared by tests/golden/make_reference_net_golden.py (which runs the reference's own network / training code over the
eager TF1 stand-in of tests/golden/tf1_shim) and by the tests that hold the oracle and the CUDA path to its output.

Nothing here knows the network: variable values are a deterministic function of the variable's TF name and shape, so
the fixture need not store the ~2 M parameters, and tensors too large to commit are summarised by three numbers
(sum, L2 norm, a fixed pseudo-random projection) that any wrong element, permutation or transposition changes.
"""
import math
import zlib

import numpy as np


def _rng(name, salt=0):
    return np.random.RandomState((zlib.crc32(name.encode()) + 7919 * salt) % (2 ** 31))


def param_value(full_name, shape, kind="weights", draw=0):
    """float32-representable 'trained-like' values: glorot-uniform weights / biases, gamma in [0.5, 1.5] (residual
    increase branches [0.05, 0.25] so the 33-block trunk is not chaotic), beta in [-0.3, 0.3].  `draw` selects another
    independent set (a fixture whose default set gives a degenerate mask records the draw it used)."""
    rng = _rng(full_name if not draw else "%s#%d" % (full_name, draw))
    shape = tuple(int(s) for s in shape)
    if kind == "gamma":
        lo, hi = (0.05, 0.25) if "increase_bn" in full_name else (0.5, 1.5)
        v = rng.uniform(lo, hi, shape)
    elif kind == "beta":
        v = rng.uniform(-0.3, 0.3, shape)
    elif kind == "moving_mean":
        v = np.zeros(shape)
    elif kind == "moving_variance":
        v = np.ones(shape)
    else:
        if len(shape) == 4:
            fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
        elif len(shape) == 2:
            fan_in, fan_out = shape
        else:
            fan_in = fan_out = shape[0]
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        v = rng.uniform(-lim, lim, shape)
    return np.ascontiguousarray(v, dtype=np.float32)


def kind_of(full_name):
    leaf = full_name.rsplit("/", 1)[-1]
    return leaf if leaf in ("gamma", "beta", "moving_mean", "moving_variance") else "weights"


def summary(name, a):
    """(sum, L2 norm, projection on a fixed N(0,1) vector seeded by the tensor's name), float64."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    p = _rng(name, salt=1).standard_normal(a.size)
    return np.array([a.sum(), math.sqrt(float((a * a).sum())), float((a * p).sum())], dtype=np.float64)


def click_map(size, where, sigma=30.0):
    """exp(-4 ln2 d^2 / sigma^2) rounded to float32 (the generator's own input data, not a parity statement)."""
    y = np.arange(size[0], dtype=np.float64)[:, None]
    x = np.arange(size[1], dtype=np.float64)[None, :]
    return np.exp(-4 * np.log(2) * ((x - where[1]) ** 2 + (y - where[0]) ** 2) / sigma ** 2).astype(np.float32)


def make_inputs(batch, size, seed, n_label_values=2, num_classes=21, ratio=8, blobs=True):
    """A synthetic batch shaped like Data.next_batch_train(): [B,S,S,4] float32 (image / 255 + click map),
    [B,S/r,S/r,1] labels with values in range(n_label_values), [B] class ids."""
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, size=(batch, size, size, 3)).astype(np.float32) / np.float32(255)
    data = np.zeros((batch, size, size, 4), dtype=np.float32)
    data[..., :3] = img
    ls = size // ratio
    lab = np.zeros((batch, ls, ls, 1), dtype=np.int64)
    for b in range(batch):
        cy, cx = rng.randint(size // 4, 3 * size // 4, size=2)
        data[b, :, :, 3] = click_map((size, size), (cy, cx))
        yy, xx = np.mgrid[0:ls, 0:ls]
        d = np.sqrt((yy - cy / ratio) ** 2 + (xx - cx / ratio) ** 2)
        for v in range(1, n_label_values):                 # nested discs: inner = highest label value
            lab[b, d < (ls / 3.0) * (n_label_values - v) / (n_label_values - 1), 0] = v
    cls = rng.randint(0, num_classes, size=(batch,)).astype(np.int64)
    return data, lab, cls

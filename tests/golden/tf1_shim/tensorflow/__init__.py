"""A minimal EAGER stand-in for the TensorFlow 1.x API surface the reference's network / training scripts touch.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_reference_net_golden.py in the build container, never shipped,
never imported by the product, the tests or bench.py).

Why it exists: the reference's arithmetic lives in TensorFlow 1.x, which cannot be installed here, so the reference
scripts cannot be run as they are.  What CAN be run is the reference's own *Python*: its `@layer` network builders
(`PSPNet.setup`, `LinkNet.build`, ...) and its `Train.build_net` / `cal_loss` methods, i.e. everything that decides
WHICH op is applied to WHAT with WHICH arguments, in WHICH order, under WHICH variable name -- layer wiring, paddings,
strides, dilation rates, bias / ReLU flags, the ReLU hidden inside `multiply`, concat order, label resizing, loss
weights, the learning-rate formula, the `var_list` of the optimizer.  This package makes `import tensorflow as tf`
resolve to an eager evaluator: every `tf.*` call the reference makes computes its value immediately on float64
torch-CPU tensors (autograd supplies `minimize`), so importing the UNMODIFIED reference modules from /root/reference
and calling their own methods yields numbers produced by the reference's own control flow.

What it does NOT prove: that each primitive below equals TensorFlow's kernel.  Every primitive is a restatement of the
published TF1 op semantics (NHWC, HWIO, 'SAME' = ceil(n/s) outputs with the odd padding element at the end,
batch-norm with the biased batch variance, align_corners bilinear, ...), written independently of
oracle/basi_oracle.py (different formulation: tap-sum convolutions and window stacks in NHWC instead of
torch.nn.functional conv2d / pooling in NCHW), so the comparison oracle <-> shim-run reference cross-checks two
restatements of the primitives and pins the network STRUCTURE to the reference's own code.
"""
from __future__ import annotations

import contextlib
import math
from collections import OrderedDict

import numpy as np
import torch

DT = torch.float64

float32 = "float32"
float64 = "float64"
int32 = "int32"
int64 = "int64"
uint8 = "uint8"
bool = "bool"  # noqa: A001  (tf.bool)


class _Anything(object):
    """Inert placeholder for API the reference's files only MENTION (decorators, initialisers, flags of the vendored
    slim model zoo that `from nets import nets_factory` imports): callable, attribute-able, usable as a decorator.
    It carries no value: if one ever reached an op of the path, that op would fail on it."""

    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and callable(a[0]) and not isinstance(a[0], _Anything):
            return a[0]                                    # used as a decorator
        return _Anything(self._name + "()")

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything(self._name + "." + n)

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ----------------------------------------------------------------------------------------------------------------
# shapes / tensors
# ----------------------------------------------------------------------------------------------------------------
class Dimension(object):
    def __init__(self, v):
        self.value = None if v is None else int(v)

    def __int__(self):
        return self.value

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return "Dimension(%r)" % self.value

    def _b(op):
        def f(self, o):
            return op(self.value, int(o))
        return f

    __mul__ = _b(lambda a, b: a * b)
    __rmul__ = __mul__
    __floordiv__ = _b(lambda a, b: a // b)
    __add__ = _b(lambda a, b: a + b)
    __sub__ = _b(lambda a, b: a - b)


class TensorShape(object):
    def __init__(self, dims):
        self.dims = [d if isinstance(d, Dimension) else Dimension(d) for d in dims]

    @property
    def ndims(self):
        return len(self.dims)

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(self.dims)

    def __getitem__(self, i):
        return TensorShape(self.dims[i]) if isinstance(i, slice) else self.dims[i]

    def as_list(self):
        return [d.value for d in self.dims]


class Tensor(object):
    """An eagerly evaluated tensor.  `.t` is the torch value (NHWC for images)."""

    def __init__(self, t, name=None, dtype=None):
        self.t = t
        self.name = name
        self._dtype = dtype

    def get_shape(self):
        return TensorShape(list(self.t.shape))

    @property
    def shape(self):
        return self.get_shape()

    @property
    def dtype(self):
        return self._dtype

    def __getitem__(self, idx):
        return Tensor(self.t[idx])

    def _bin(op):
        def f(self, o):
            return Tensor(op(self.t, _val(o)))
        return f

    def _rbin(op):
        def f(self, o):
            return Tensor(op(_val(o), self.t))
        return f

    __add__ = _bin(lambda a, b: a + b)
    __radd__ = _rbin(lambda a, b: a + b)
    __sub__ = _bin(lambda a, b: a - b)
    __rsub__ = _rbin(lambda a, b: a - b)
    __mul__ = _bin(lambda a, b: a * b)
    __rmul__ = _rbin(lambda a, b: a * b)
    __truediv__ = _bin(lambda a, b: a / b)
    __rtruediv__ = _rbin(lambda a, b: a / b)
    __div__ = __truediv__
    __neg__ = lambda self: Tensor(-self.t)      # noqa: E731
    __gt__ = _bin(lambda a, b: a > b)
    __ge__ = _bin(lambda a, b: a >= b)
    __lt__ = _bin(lambda a, b: a < b)
    __le__ = _bin(lambda a, b: a <= b)


class Variable(Tensor):
    def __init__(self, t, name, trainable):
        Tensor.__init__(self, t, name + ":0")
        self.full_name = name
        self.trainable = trainable
        self.op = type("Op", (), {"name": name})()


def _val(x):
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, Dimension):
        return x.value
    return x


def _wrap(t, name=None):
    return Tensor(t, name)


# ----------------------------------------------------------------------------------------------------------------
# graph state: variables, scopes, feeds, op trace
# ----------------------------------------------------------------------------------------------------------------
class _State(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.variables = OrderedDict()      # full name -> Variable
        self.scope = []                     # variable-scope name stack
        self.feeds = []                     # values handed out by tf.placeholder, in creation order
        self.placeholders = []
        self.provider = None                # (full_name, shape, kind) -> numpy array
        self.trace = []                     # (op, attrs) of every structural call, in execution order


_S = _State()


def shim_reset(provider, feeds):
    """Start a fresh 'graph': `provider(full_name, shape, kind)` supplies variable values (kind in weights / gamma /
    beta / moving_mean / moving_variance), `feeds` are the placeholder values in creation order (a list, or a callable
    (index, dtype, shape) -> value for scripts that create their reader before their placeholders)."""
    _S.reset()
    _S.provider = provider
    _S.feeds = feeds if callable(feeds) else list(feeds)


def shim_state():
    return _S


def _trace(op, **attrs):
    _S.trace.append((op, attrs))


class VariableScope(object):
    def __init__(self, name):
        self.name = name
        self.original_name_scope = name + "/"


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, values=None, reuse=None, **_):
    saved = list(_S.scope)
    if isinstance(name_or_scope, VariableScope):
        _S.scope = [p for p in name_or_scope.name.split("/") if p]
    else:
        n = name_or_scope if name_or_scope is not None else default_name
        if n:
            _S.scope = saved + [n]
    try:
        yield VariableScope("/".join(_S.scope))
    finally:
        _S.scope = saved


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    yield name or default_name


def get_variable_scope():
    return VariableScope("/".join(_S.scope))


def _kind_of(name):
    return name if name in ("gamma", "beta", "moving_mean", "moving_variance") else "weights"


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **_):
    full = "/".join(_S.scope + [name])
    if full in _S.variables:
        raise ValueError("Variable %s already exists (the reference never reuses variables)" % full)
    shape = tuple(int(_val(s)) for s in shape)
    v = np.asarray(_S.provider(full, shape, _kind_of(name)), dtype=np.float64)
    assert v.shape == shape, (full, v.shape, shape)
    t = torch.tensor(v, dtype=DT, requires_grad=True)
    var = Variable(t, full, trainable)
    _S.variables[full] = var
    return var


def trainable_variables():
    return [v for v in _S.variables.values() if v.trainable]


def global_variables():
    return list(_S.variables.values())


def placeholder(dtype=None, shape=None, name=None):
    i = len(_S.placeholders)
    v = _S.feeds(i, dtype, shape) if callable(_S.feeds) else _S.feeds[i]
    if dtype in (int32, int64, uint8):
        t = torch.as_tensor(np.asarray(v)).to(torch.int64)
    else:
        t = torch.as_tensor(np.asarray(v, dtype=np.float64))
    if shape is not None:
        want = [None if s is None else int(_val(s)) for s in shape]
        assert len(want) == t.dim() and all(w is None or w == g for w, g in zip(want, t.shape)), \
            "feed %d has shape %s, placeholder wants %s" % (len(_S.placeholders), tuple(t.shape), want)
    p = Tensor(t, name, dtype)
    _S.placeholders.append(p)
    return p


def placeholder_with_default(input, shape=None, name=None):      # noqa: A002
    return input


def constant(value, dtype=None, shape=None, name=None):
    return Tensor(torch.as_tensor(np.asarray(value, dtype=np.float64)), name)


# ----------------------------------------------------------------------------------------------------------------
# array ops
# ----------------------------------------------------------------------------------------------------------------
class _ShapeValue(list):
    """Result of tf.shape(): a list of python ints (eager), sliceable like the int32 tensor."""

    def __getitem__(self, i):
        r = list.__getitem__(self, i)
        return _ShapeValue(r) if isinstance(i, slice) else r


def shape(input, name=None):      # noqa: A002
    return _ShapeValue(int(s) for s in _val(input).shape)


def stack(values, axis=0, name=None):
    if all(not isinstance(v, Tensor) for v in values):
        return _ShapeValue(int(_val(v)) for v in values)
    return Tensor(torch.stack([_val(v) for v in values], dim=axis))


def _size2(size):
    return [int(_val(s)) for s in size]


def reshape(tensor, shape, name=None):      # noqa: A002
    return Tensor(_val(tensor).reshape([int(_val(s)) for s in shape]), name)


def squeeze(input, axis=None, name=None, squeeze_dims=None):      # noqa: A002
    ax = axis if axis is not None else squeeze_dims
    t = _val(input)
    for a in sorted(ax, reverse=True):
        assert t.shape[a] == 1, "tf.squeeze: dimension %d has size %d" % (a, t.shape[a])
        t = t.squeeze(a)
    _trace("squeeze", axis=list(ax))
    return Tensor(t, name)


def expand_dims(input, axis=None, name=None, dim=None):      # noqa: A002
    return Tensor(_val(input).unsqueeze(axis if axis is not None else dim), name)


def concat(values, axis, name=None):
    _trace("concat", axis=axis, n=len(values), channels=[int(_val(v).shape[-1]) for v in values])
    return Tensor(torch.cat([_val(v) for v in values], dim=axis), name)


def split(value, num_or_size_splits, axis=0, num=None, name="split"):
    t = _val(value)
    if isinstance(num_or_size_splits, int):
        assert t.shape[axis] % num_or_size_splits == 0
        parts = torch.split(t, t.shape[axis] // num_or_size_splits, dim=axis)
    else:
        parts = torch.split(t, list(num_or_size_splits), dim=axis)
    return [Tensor(p) for p in parts]


def pad(tensor, paddings, mode="CONSTANT", name=None, constant_values=0):
    t = _val(tensor)
    p = np.asarray(paddings).astype(int)
    assert p.shape == (t.dim(), 2)
    out = t.new_zeros([t.shape[i] + int(p[i, 0]) + int(p[i, 1]) for i in range(t.dim())])
    idx = tuple(slice(int(p[i, 0]), int(p[i, 0]) + t.shape[i]) for i in range(t.dim()))
    out[idx] = t
    _trace("pad", paddings=p.tolist())
    return Tensor(out, name)


def cast(x, dtype, name=None):
    t = _val(x)
    if dtype in (int32, int64, uint8):
        return Tensor(t.to(torch.int64), name, dtype)         # float -> int truncates toward zero, like tf.cast
    if dtype == bool:
        return Tensor(t != 0, name, dtype)
    return Tensor(t.to(DT), name, dtype)


def greater(x, y, name=None):
    return Tensor(_val(x) > _val(y), name)


def argmax(input, axis=None, name=None, dimension=None, output_type=int64):      # noqa: A002
    ax = axis if axis is not None else dimension
    return Tensor(torch.argmax(_val(input), dim=ax), name)       # first maximal index, like TF


def one_hot(indices, depth, on_value=1.0, off_value=0.0, axis=-1, dtype=None, name=None):
    i = _val(indices).to(torch.int64)
    out = torch.zeros(tuple(i.shape) + (int(depth),), dtype=DT)
    out.scatter_(-1, i.unsqueeze(-1), 1.0)
    return Tensor(out * (on_value - off_value) + off_value, name)


def where(condition, x=None, y=None, name=None):
    return Tensor(torch.where(_val(condition), _val(x), _val(y)), name)


def zeros_like(tensor, dtype=None, name=None, optimize=True):
    return Tensor(torch.zeros_like(_val(tensor)), name)


def ones_like(tensor, dtype=None, name=None, optimize=True):
    return Tensor(torch.ones_like(_val(tensor)), name)


# ----------------------------------------------------------------------------------------------------------------
# math ops
# ----------------------------------------------------------------------------------------------------------------
def multiply(x, y, name=None):
    _trace("multiply", channels=[int(_val(x).shape[-1]), int(_val(y).shape[-1])])
    return Tensor(_val(x) * _val(y), name)


def add(x, y, name=None):
    return Tensor(_val(x) + _val(y), name)


def add_n(inputs, name=None):
    _trace("add_n", n=len(inputs))
    t = _val(inputs[0])
    for i in inputs[1:]:
        t = t + _val(i)
    return Tensor(t, name)


def scalar_mul(scalar, x):
    return Tensor(_val(scalar) * _val(x))


def pow(x, y, name=None):      # noqa: A001
    return Tensor(torch.pow(torch.as_tensor(_val(x), dtype=DT), _val(y)), name)


def reduce_mean(input_tensor, axis=None, keepdims=False, name=None, **_):
    t = _val(input_tensor)
    return Tensor(t.mean() if axis is None else t.mean(dim=axis, keepdim=keepdims), name)


def reduce_sum(input_tensor, axis=None, keepdims=False, name=None, **_):
    t = _val(input_tensor)
    return Tensor(t.sum() if axis is None else t.sum(dim=axis, keepdim=keepdims), name)


def sigmoid(x, name=None):
    return nn.sigmoid(x, name)


def equal(x, y, name=None):
    return Tensor(_val(x) == _val(y), name)


# ----------------------------------------------------------------------------------------------------------------
# tf.nn
# ----------------------------------------------------------------------------------------------------------------
def _same_pad(n, k_eff, s):
    """TF 'SAME': ceil(n / s) outputs; total padding split with the odd element at the END."""
    out = -(-n // s)
    total = max((out - 1) * s + k_eff - n, 0)
    return total // 2, total - total // 2


def _pad_hw(t, ph, pw, value=0.0):
    if ph == (0, 0) and pw == (0, 0):
        return t
    b, h, w, c = t.shape
    out = t.new_full((b, h + ph[0] + ph[1], w + pw[0] + pw[1], c), value)
    out[:, ph[0]:ph[0] + h, pw[0]:pw[0] + w, :] = t
    return out


def _conv_nhwc(x, w, sh, sw, padding, dh=1, dw=1):
    """Cross-correlation as a sum over filter taps of (strided window) @ w[i, j]: NHWC input, HWIO filter."""
    kh, kw, ci, co = w.shape
    assert x.shape[-1] == ci, "conv: input has %d channels, filter expects %d" % (x.shape[-1], ci)
    keh, kew = (kh - 1) * dh + 1, (kw - 1) * dw + 1
    if padding == "SAME":
        x = _pad_hw(x, _same_pad(x.shape[1], keh, sh), _same_pad(x.shape[2], kew, sw))
    else:
        assert padding == "VALID", padding
    ho = (x.shape[1] - keh) // sh + 1
    wo = (x.shape[2] - kew) // sw + 1
    y = None
    for i in range(kh):
        for j in range(kw):
            win = x[:, i * dh: i * dh + (ho - 1) * sh + 1: sh, j * dw: j * dw + (wo - 1) * sw + 1: sw, :]
            term = torch.matmul(win, w[i, j])
            y = term if y is None else y + term
    return y


class _NN(object):
    @staticmethod
    def conv2d(input, filter=None, strides=None, padding=None, use_cudnn_on_gpu=True, data_format="NHWC",      # noqa: A002
               dilations=None, name=None, filters=None):
        w = _val(filter if filter is not None else filters)
        assert data_format == "NHWC" and strides[0] == 1 and strides[3] == 1
        _trace("conv2d", k=[int(w.shape[0]), int(w.shape[1])], cin=int(w.shape[2]), cout=int(w.shape[3]),
               strides=[int(strides[1]), int(strides[2])], padding=padding, rate=1)
        return Tensor(_conv_nhwc(_val(input), w, int(strides[1]), int(strides[2]), padding), name)

    @staticmethod
    def atrous_conv2d(value, filters, rate, padding, name=None):
        w = _val(filters)
        _trace("conv2d", k=[int(w.shape[0]), int(w.shape[1])], cin=int(w.shape[2]), cout=int(w.shape[3]),
               strides=[1, 1], padding=padding, rate=int(rate))
        return Tensor(_conv_nhwc(_val(value), w, 1, 1, padding, int(rate), int(rate)), name)

    @staticmethod
    def bias_add(value, bias, data_format=None, name=None):
        _trace("bias_add", c=int(_val(bias).shape[0]))
        return Tensor(_val(value) + _val(bias), name)

    @staticmethod
    def relu(features, name=None):
        _trace("relu")
        return Tensor(torch.clamp(_val(features), min=0.0), name)

    @staticmethod
    def sigmoid(x, name=None):
        _trace("sigmoid")
        return Tensor(torch.sigmoid(_val(x)), name)

    @staticmethod
    def softmax(logits, axis=None, name=None, dim=None):
        if isinstance(axis, str):          # the reference passes the layer name positionally (Network.softmax)
            axis = None
        a = -1 if axis is None and dim is None else (axis if axis is not None else dim)
        _trace("softmax", axis=a)
        return Tensor(torch.softmax(_val(logits), dim=a), name)

    @staticmethod
    def _pool(value, ksize, strides, padding, kind):
        x = _val(value)
        assert ksize[0] == 1 and ksize[3] == 1 and strides[0] == 1 and strides[3] == 1
        kh, kw, sh, sw = int(ksize[1]), int(ksize[2]), int(strides[1]), int(strides[2])
        _trace(kind, k=[kh, kw], strides=[sh, sw], padding=padding)
        if padding == "SAME":
            ph, pw = _same_pad(x.shape[1], kh, sh), _same_pad(x.shape[2], kw, sw)
        else:
            ph = pw = (0, 0)
        if kind == "max_pool":
            xp = _pad_hw(x, ph, pw, -math.inf)
        else:
            xp = _pad_hw(x, ph, pw, 0.0)
            cnt = _pad_hw(torch.ones_like(x[..., :1]), ph, pw, 0.0)      # SAME avg-pool divides by the valid count
        ho, wo = (xp.shape[1] - kh) // sh + 1, (xp.shape[2] - kw) // sw + 1

        def windows(t):
            return torch.stack([t[:, i: i + (ho - 1) * sh + 1: sh, j: j + (wo - 1) * sw + 1: sw, :]
                                for i in range(kh) for j in range(kw)], dim=0)
        if kind == "max_pool":
            return windows(xp).max(dim=0).values
        return windows(xp).sum(dim=0) / windows(cnt).sum(dim=0)

    @staticmethod
    def max_pool(value, ksize, strides, padding, data_format="NHWC", name=None):
        return Tensor(_NN._pool(value, ksize, strides, padding, "max_pool"), name)

    @staticmethod
    def avg_pool(value, ksize, strides, padding, data_format="NHWC", name=None):
        return Tensor(_NN._pool(value, ksize, strides, padding, "avg_pool"), name)

    @staticmethod
    def xw_plus_b(x, weights, biases, name=None):
        _trace("xw_plus_b", shape=[int(s) for s in _val(weights).shape])
        return Tensor(torch.matmul(_val(x), _val(weights)) + _val(biases), name)

    @staticmethod
    def relu_layer(x, weights, biases, name=None):
        _trace("relu_layer", shape=[int(s) for s in _val(weights).shape])
        return Tensor(torch.clamp(torch.matmul(_val(x), _val(weights)) + _val(biases), min=0.0), name)

    @staticmethod
    def dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
        raise NotImplementedError("dropout is defined by the reference's Network class but never used on this path")

    @staticmethod
    def local_response_normalization(*a, **k):
        raise NotImplementedError("lrn is defined by the reference's Network class but never used on this path")

    @staticmethod
    def batch_normalization(*a, **k):
        raise NotImplementedError("only inside a dead docstring of the reference")

    @staticmethod
    def weighted_cross_entropy_with_logits(targets=None, logits=None, pos_weight=None, name=None, labels=None):
        """TF1 definition: targets * -log(sigmoid(x)) * pos_weight + (1 - targets) * -log(1 - sigmoid(x))."""
        z = _val(targets if targets is not None else labels).to(DT)
        x = _val(logits)
        _trace("weighted_cross_entropy_with_logits", pos_weight=float(pos_weight))
        # softplus(-x) = -log(sigmoid(x)), softplus(x) = -log(1 - sigmoid(x))
        sp = torch.nn.functional.softplus
        return Tensor(z * float(pos_weight) * sp(-x) + (1.0 - z) * sp(x), name)

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
        x = _val(logits)
        lab = _val(labels).to(torch.int64)
        _trace("sparse_softmax_cross_entropy_with_logits", classes=int(x.shape[-1]))
        lse = torch.logsumexp(x, dim=-1)
        picked = torch.gather(x, -1, lab.unsqueeze(-1)).squeeze(-1)
        return Tensor(lse - picked, name)

    @staticmethod
    def softmax_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, dim=-1, name=None):
        x = _val(logits)
        _trace("softmax_cross_entropy_with_logits", classes=int(x.shape[-1]))
        return Tensor(-(torch.log_softmax(x, dim=dim) * _val(labels).to(DT)).sum(dim=dim), name)


    def __getattr__(self, name):          # e.g. tf.nn.relu6 in a model-zoo file that is imported but never built
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything("tf.nn." + name)


nn = _NN()


# ----------------------------------------------------------------------------------------------------------------
# tf.layers
# ----------------------------------------------------------------------------------------------------------------
class _Layers(object):
    @staticmethod
    def batch_normalization(inputs, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, training=False,
                            trainable=True, name=None, reuse=None, **_):
        """tf.layers.batch_normalization: variables gamma / beta / moving_mean / moving_variance under scope `name`;
        training=True normalises with the batch mean and the BIASED batch variance over N, H, W (tf.nn.moments)."""
        x = _val(inputs)
        c = x.shape[-1]
        with variable_scope(name or "batch_normalization"):
            gamma = get_variable("gamma", [c], trainable=trainable) if scale else None
            beta = get_variable("beta", [c], trainable=trainable) if center else None
            mm = get_variable("moving_mean", [c], trainable=False)
            mv = get_variable("moving_variance", [c], trainable=False)
        _trace("batch_normalization", c=int(c), momentum=float(momentum), epsilon=float(epsilon), training=training
               if isinstance(training, (int, type(None))) else "tensor")
        red = tuple(range(x.dim() - 1))
        if training:
            mean = x.mean(dim=red, keepdim=True)
            var = ((x - mean) * (x - mean)).mean(dim=red, keepdim=True)
        else:
            mean, var = mm.t.detach(), mv.t.detach()
        y = (x - mean) / torch.sqrt(var + epsilon)
        if gamma is not None:
            y = y * gamma.t
        if beta is not None:
            y = y + beta.t
        return Tensor(y, name)


layers = _Layers()


# ----------------------------------------------------------------------------------------------------------------
# tf.image
# ----------------------------------------------------------------------------------------------------------------
class _Image(object):
    @staticmethod
    def resize_bilinear(images, size, align_corners=False, name=None):
        x = _val(images)
        ho, wo = _size2(size)
        hi, wi = x.shape[1], x.shape[2]
        _trace("resize_bilinear", align_corners=align_corners)

        def axis_weights(n_in, n_out):
            if align_corners and n_out > 1:
                scale = (n_in - 1) / float(n_out - 1)
            else:
                scale = n_in / float(n_out)
            m = torch.zeros(n_out, n_in, dtype=DT)
            for o in range(n_out):
                src = o * scale                          # TF1 legacy mapping (no half-pixel centres)
                lo = min(int(math.floor(src)), n_in - 1)
                hi_ = min(lo + 1, n_in - 1)
                f = src - lo
                m[o, lo] += 1.0 - f
                m[o, hi_] += f
            return m
        mh, mw = axis_weights(hi, ho), axis_weights(wi, wo)
        y = torch.einsum("oh,bhwc->bowc", mh, x)
        y = torch.einsum("pw,bowc->bopc", mw, y)
        return Tensor(y, name)

    @staticmethod
    def resize_nearest_neighbor(images, size, align_corners=False, name=None):
        x = _val(images)
        ho, wo = _size2(size)
        hi, wi = x.shape[1], x.shape[2]
        _trace("resize_nearest_neighbor", align_corners=align_corners)

        def idx(n_in, n_out):
            if align_corners and n_out > 1:
                scale = (n_in - 1) / float(n_out - 1)
                return [min(int(round(o * scale)), n_in - 1) for o in range(n_out)]
            # TF's kernel works in float32: floorf(out * (float(in) / float(out))) -- 224 -> 110 maps row 55 to
            # 112 (float64 arithmetic would say 111)
            scale = np.float32(n_in) / np.float32(n_out)
            return [min(int(np.floor(np.float32(o) * scale)), n_in - 1) for o in range(n_out)]
        ih = torch.tensor(idx(hi, ho), dtype=torch.int64)
        iw = torch.tensor(idx(wi, wo), dtype=torch.int64)
        return Tensor(x.index_select(1, ih).index_select(2, iw), name)


image = _Image()


# ----------------------------------------------------------------------------------------------------------------
# tf.train
# ----------------------------------------------------------------------------------------------------------------
class _TrainOp(object):
    """What `minimize` returns: the gradients and the SGD-updated values (the eager stand-in for running the op)."""

    def __init__(self, lr, names, grads, new_values):
        self.learning_rate = lr
        self.var_names = names
        self.grads = grads
        self.new_values = new_values


class _GradientDescentOptimizer(object):
    def __init__(self, learning_rate, use_locking=False, name="GradientDescent"):
        self.lr = learning_rate

    def minimize(self, loss, global_step=None, var_list=None, **_):
        vs = list(var_list) if var_list is not None else trainable_variables()
        gs = torch.autograd.grad(_val(loss), [v.t for v in vs], retain_graph=True, allow_unused=True)
        lr = float(_val(self.lr))
        names, grads, new = [], OrderedDict(), OrderedDict()
        for v, g in zip(vs, gs):
            if g is None:                      # TF: "No gradients provided" variables are skipped by apply_gradients
                continue
            names.append(v.full_name)
            grads[v.full_name] = g.detach().numpy().copy()
            new[v.full_name] = (v.t.detach() - lr * g).numpy().copy()
        return _TrainOp(lr, names, grads, new)


class _Inert(object):
    """Saver / FileWriter / Session: control plane the scripts construct.  Session.run(fetches, feed_dict) returns the
    already (eagerly) computed values of the fetched tensors after checking that every fed value is the one its
    placeholder was evaluated with -- enough for the reference's inference scripts, which build the graph and run it
    once in the same method."""

    def __init__(self, *a, **k):
        self.graph = None
        self.runs = []

    def run(self, fetches=None, feed_dict=None, **k):
        if fetches is None:
            return None
        for ph, v in (feed_dict or {}).items():
            assert isinstance(ph, Tensor) and np.array_equal(np.asarray(v, dtype=np.float64),
                                                             ph.t.detach().numpy().astype(np.float64)), \
                "Session.run: a fed value differs from the one the stand-in evaluated the graph with"

        def value(f):
            return None if f is None else f.t.detach().numpy()
        out = [value(f) for f in fetches] if isinstance(fetches, (list, tuple)) else value(fetches)
        self.runs.append(out)
        SESSION_RUNS.append(out)
        return out

    def add_summary(self, *a, **k):      # FileWriter: TensorBoard dumps, control plane
        return None

    def __getattr__(self, name):
        raise NotImplementedError("control plane (%s): the stand-in evaluates eagerly, there is nothing to run" % name)


SESSION_RUNS = []          # what every Session.run returned, in order (read by the golden generator)


class _Train(object):
    GradientDescentOptimizer = _GradientDescentOptimizer
    Saver = _Inert

    @staticmethod
    def get_checkpoint_state(checkpoint_dir, latest_filename=None):
        return None            # "No checkpoint file found.": the variables keep the values the provider gave them


train = _Train()


# ----------------------------------------------------------------------------------------------------------------
# control plane the scripts touch in __init__ (never exercised by the golden generator)
# ----------------------------------------------------------------------------------------------------------------
class _Summary(object):
    FileWriter = _Inert

    @staticmethod
    def scalar(*a, **k):
        return None

    @staticmethod
    def image(*a, **k):
        return None

    @staticmethod
    def merge_all(*a, **k):
        return None


summary = _Summary()
Session = _Inert


def ConfigProto(*a, **k):
    return None


def GPUOptions(*a, **k):
    return None


def global_variables_initializer():
    return None


def zeros_initializer(*a, **k):
    return "zeros"


def __getattr__(name):                    # API the vendored model zoo only mentions (see _Anything)
    if name.startswith("__"):
        raise AttributeError(name)
    return _Anything("tf." + name)

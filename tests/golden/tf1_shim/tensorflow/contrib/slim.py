"""tf.contrib.slim stand-in (see ../__init__.py): exactly what slim/nets/vgg.py's vgg_arg_scope + vgg_16 and
slim/nets/nets_factory.get_network_fn use -- arg_scope, conv2d, max_pool2d, repeat, dropout, l2_regularizer,
utils.convert_collection_to_dict.  Everything else the other slim model files touch at import time resolves to an inert
placeholder (module __getattr__), so `from nets import nets_factory` imports the reference's whole model zoo unmodified.

slim.conv2d semantics restated: variables `<scope>/weights` [kh, kw, cin, cout] and `<scope>/biases` [cout],
stride 1, padding from the arg scope ('SAME' for vgg), output = activation_fn(conv + biases); the output is registered in
`outputs_collections` under the alias of its variable scope (that is how vgg_16's end_points dict is built).
slim.max_pool2d: kernel [2, 2], stride 2 (slim's default), padding 'VALID'.
"""
import contextlib

from .. import (_Anything, _trace, _val, Tensor, get_variable, get_variable_scope, nn, variable_scope)

_SCOPE_STACK = [{}]            # {function name: default kwargs}
_COLLECTIONS = {}              # collection name -> [(alias, tensor)]


def _key(f):
    return getattr(f, "_slim_name", getattr(f, "__name__", repr(f)))


@contextlib.contextmanager
def arg_scope(list_ops_or_scope, **kwargs):
    if isinstance(list_ops_or_scope, dict):                # re-entering a captured scope (nets_factory does this)
        new = {k: dict(v) for k, v in _SCOPE_STACK[-1].items()}
        for k, v in list_ops_or_scope.items():
            new.setdefault(k, {}).update(v)
    else:
        new = {k: dict(v) for k, v in _SCOPE_STACK[-1].items()}
        for f in list_ops_or_scope:
            new.setdefault(_key(f), {}).update(kwargs)
    _SCOPE_STACK.append(new)
    try:
        yield new
    finally:
        _SCOPE_STACK.pop()


def add_arg_scope(f):
    name = f.__name__

    def wrapped(*a, **k):
        merged = dict(_SCOPE_STACK[-1].get(name, {}))
        merged.update(k)
        return f(*a, **merged)
    wrapped.__name__ = name
    wrapped._slim_name = name
    return wrapped


def _collect(outputs_collections, tensor):
    if outputs_collections:
        _COLLECTIONS.setdefault(outputs_collections, []).append((get_variable_scope().name, tensor))
    return tensor


@add_arg_scope
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn=nn.relu, normalizer_fn=None,
           weights_initializer=None, weights_regularizer=None, biases_initializer="zeros", biases_regularizer=None,
           outputs_collections=None, trainable=True, scope=None, rate=1, reuse=None, **_):
    assert normalizer_fn is None and rate == 1
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    with variable_scope(scope, default_name="Conv"):
        w = get_variable("weights", [kh, kw, inputs.get_shape()[-1], num_outputs], trainable=trainable)
        out = nn.conv2d(inputs, w, [1, stride, stride, 1], padding=padding)
        if biases_initializer is not None:
            out = nn.bias_add(out, get_variable("biases", [num_outputs], trainable=trainable))
        if activation_fn is not None:
            out = activation_fn(out)
        return _collect(outputs_collections, out)


@add_arg_scope
def max_pool2d(inputs, kernel_size, stride=2, padding="VALID", outputs_collections=None, scope=None, **_):
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    with variable_scope(scope, default_name="MaxPool2D"):
        return _collect(outputs_collections, nn.max_pool(inputs, [1, kh, kw, 1], [1, stride, stride, 1], padding))


@add_arg_scope
def fully_connected(*a, **k):
    raise NotImplementedError("not on the path (vgg_16 uses convolutions for fc6 / fc7)")


@add_arg_scope
def dropout(inputs, keep_prob=0.5, is_training=True, scope=None, **_):
    """vgg_16 applies dropout to fc6 only; the reference's LinkNet never reads anything downstream of pool5, so the
    random mask cannot reach a value on the path.  Identity, traced so the golden generator can assert exactly that."""
    _trace("dropout", keep_prob=float(keep_prob), is_training=bool(is_training))
    return Tensor(_val(inputs) * 1.0)


def repeat(inputs, repetitions, layer, *args, **kwargs):
    scope = kwargs.pop("scope", None)
    with variable_scope(scope, default_name="Repeat"):
        out = inputs
        for i in range(repetitions):
            kwargs["scope"] = "%s_%d" % (scope, i + 1)
            out = layer(out, *args, **kwargs)
        return out


def l2_regularizer(scale, scope=None):
    return ("l2_regularizer", scale)      # collected by TF into REGULARIZATION_LOSSES; the reference never adds them


class utils(object):
    @staticmethod
    def convert_collection_to_dict(collection, clear_collection=False):
        return dict(_COLLECTIONS.get(collection, []))


def shim_reset_collections():
    _COLLECTIONS.clear()
    del _SCOPE_STACK[1:]


def __getattr__(name):                    # anything else the model zoo mentions at import time
    return _Anything("slim." + name)

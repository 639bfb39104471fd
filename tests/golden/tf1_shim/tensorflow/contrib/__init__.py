"""tf.contrib stand-in (see ../__init__.py): `metrics.accuracy`, the handful of slim layers vgg_16 uses, inert
placeholders for everything else the vendored model zoo mentions."""
from .. import _Anything
from . import metrics, slim  # noqa: F401


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Anything("tf.contrib." + name)

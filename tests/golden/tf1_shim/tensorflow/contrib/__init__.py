"""tf.contrib stand-in (see ../__init__.py): only `metrics.accuracy` and the handful of slim layers vgg_16 uses."""
from . import metrics  # noqa: F401

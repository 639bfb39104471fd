"""tf.contrib.metrics.accuracy: the fraction of equal elements."""
import torch

from .. import DT, Tensor, _val


def accuracy(predictions, labels, weights=None, name=None):
    p, l = _val(predictions), _val(labels)
    assert p.shape == l.shape, (p.shape, l.shape)
    return Tensor((p == l).to(DT).mean(), name)

"""Feedable values: what ``tf.placeholder`` becomes without a graph.

The reference feeds minibatches and the stepsize into ``session.run`` through
``feed_dict = {placeholder: value}`` (pysgmcmc/samplers/base_classes.py:124-197,
pysgmcmc/data_batches.py:125-129).  Here a `Placeholder` is a named slot: the
sampler stores each fed value in ``placeholder.value`` right before it evaluates the
cost function, and cost functions read it from there.
"""
import numpy as np
import torch


class Placeholder(object):
    def __init__(self, name=None, dtype=None, shape=None):
        self.name = name
        self.dtype = dtype
        self.shape = shape
        self.value = None

    def tensor(self, device, dtype=None):
        """The fed value as a tensor on `device`."""
        v = self.value
        if v is None:
            raise ValueError("placeholder %r was evaluated before a value was fed" % (self.name,))
        if not isinstance(v, torch.Tensor):
            v = torch.as_tensor(np.asarray(v))
        return v.to(device=device, dtype=dtype if dtype is not None else self.dtype or v.dtype)

    def __float__(self):
        return float(self.value)

    def __repr__(self):
        return "Placeholder(name=%r)" % (self.name,)


def placeholder(dtype=None, shape=None, name=None):
    """Signature-compatible stand-in for ``tf.placeholder(dtype, shape, name)``."""
    return Placeholder(name=name, dtype=dtype, shape=shape)


def feed(feed_dict):
    """Store every ``{Placeholder: value}`` entry of `feed_dict` in its placeholder."""
    for key, value in feed_dict.items():
        if isinstance(key, Placeholder):
            key.value = value

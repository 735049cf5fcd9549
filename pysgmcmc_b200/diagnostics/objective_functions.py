"""Test densities of the sampler hot path -- same names and semantics as
pysgmcmc/diagnostics/objective_functions.py:7-102 (banana, gmm1-3,
to_negative_log_likelihood, sinc).

Each log likelihood accepts torch tensors (differentiable; with a leading chain
axis the result is one value per chain), NumPy arrays or Python numbers.  The
functions carry a ``native_target`` tag: when a sampler receives
``to_negative_log_likelihood(banana_log_likelihood)`` (or a gmm) as its cost
function it runs the fused CUDA kernel K6 (csrc/target_chains.cu) instead of
differentiating the Python callable.  The HPOLIB regression functions of the
reference (:107-315) are not part of the sampler path and are not provided.
"""
import functools
import math

import numpy as np
import torch


def to_negative_log_likelihood(log_likelihood_function):
    """Decorator: log likelihood -> negative log likelihood (objective_functions.py:7-45).

    >>> import numpy as np
    >>> log_likelihood = lambda a, b: np.log(a + b)
    >>> negative_log_likelihood = to_negative_log_likelihood(log_likelihood)
    >>> bool(np.allclose(-log_likelihood(4, 5), negative_log_likelihood(4, 5)))
    True
    >>> log_likelihood.__name__ == negative_log_likelihood.__name__
    True
    """
    @functools.wraps(log_likelihood_function)
    def negative_log_likelihood(*args, **kwargs):
        return -log_likelihood_function(*args, **kwargs)
    tag = getattr(log_likelihood_function, "native_target", None)
    if tag is not None:
        negative_log_likelihood.native_target = (tag[0], -tag[1])
    return negative_log_likelihood


def _native(name):
    def mark(fn):
        fn.native_target = (name, +1)     # (+1: log likelihood, -1: cost = negative log likelihood)
        return fn
    return mark


@_native("banana")
def banana_log_likelihood(x):
    """objective_functions.py:49-59.

    >>> float(banana_log_likelihood((0, 10)))
    -0.0
    """
    return -0.5 * (0.01 * x[0] ** 2 + (x[1] + 0.1 * x[0] ** 2 - 10) ** 2)


def gaussian_mixture_model_log_likelihood(x, mu=(-5, 0, 5), var=(1., 1., 1.),
                                          weights=(1. / 3., 1. / 3., 1. / 3.)):
    """objective_functions.py:62-85 (1-d only)."""
    assert len(mu) == len(var) == len(weights)

    if isinstance(x, (list, tuple)):
        assert(len(x) == 1)
        x = x[0]

    if isinstance(x, torch.Tensor):
        comps = [math.log(weights[i]) + (-0.5 * math.log(2.0 * math.pi * var[i])
                                        - 0.5 * ((x - mu[i]) ** 2) / var[i])
                 for i in range(len(mu))]
        return torch.logsumexp(torch.stack(comps, dim=0), dim=0)

    def normldf(x, mu, var):
        return -0.5 * np.log(2.0 * np.pi * var) - 0.5 * ((x - mu) ** 2) / var

    from scipy.special import logsumexp
    return logsumexp([np.log(weights[i]) + normldf(x, mu[i], var[i]) for i in range(len(mu))], axis=0)


@_native("gmm1")
def gmm1_log_likelihood(x):
    return gaussian_mixture_model_log_likelihood(x)


@_native("gmm2")
def gmm2_log_likelihood(x):
    return gaussian_mixture_model_log_likelihood(x, var=[1. / 0.5, 0.5, 1. / 0.5])


@_native("gmm3")
def gmm3_log_likelihood(x):
    return gaussian_mixture_model_log_likelihood(x, var=[1. / 0.3, 0.3, 1. / 0.3])


def sinc(x):
    """objective_functions.py:101-102."""
    return np.sinc(x * 10 - 5).sum(axis=1)

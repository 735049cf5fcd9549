"""Model base class and the normalisation helpers the BNN path uses.

Public names and behaviour follow pysgmcmc/models/base_model.py (BaseModel :5-108, the four
normalisation functions :109-137) so that model code written against the reference keeps
working; the implementation is this engine's own: the array-shape contracts live in one helper
(`_require`) that both decorators share, and the affine normalisations are two instances of one
transform.
"""
import abc
import functools

import numpy as np

__all__ = ("BaseModel", "zero_one_normalization", "zero_one_unnormalization",
           "zero_mean_unit_var_normalization", "zero_mean_unit_var_unnormalization")


def _require(condition):
    # the reference signals malformed inputs with bare asserts (base_model.py:68-70,76);
    # pysgmcmc/tests/bayesian_neural_network/test_invalid_inputs.py expects AssertionError
    if not condition:
        raise AssertionError


def _shape_checked(n_array_args):
    """Decorator factory: the first `n_array_args` positional arguments are `X` (2-d) and,
    if two, `y` (1-d, one entry per row of X)."""
    def decorate(method):
        @functools.wraps(method)
        def checked(self, X, *rest, **kwargs):
            if n_array_args == 2:
                _require(len(rest) >= 1 or "y" in kwargs)
                y = rest[0] if rest else kwargs["y"]
                _require(X.shape[0] == y.shape[0])
                _require(X.ndim == 2)
                _require(y.ndim == 1)
            else:
                _require(X.ndim == 2)
            return method(self, X, *rest, **kwargs)
        return checked
    return decorate


class BaseModel(abc.ABC):
    """Interface of a regression model: ``train(X, y)``, ``predict(X_test) -> (mean, var)``,
    ``update(X, y)`` (retrain on old + new data) and ``get_incumbent()``."""

    #: decorators for subclasses' ``train`` / ``predict`` (used as ``@BaseModel._check_shapes_train``)
    _check_shapes_train = staticmethod(_shape_checked(2))
    _check_shapes_predict = staticmethod(_shape_checked(1))

    def __init__(self):
        self.X, self.y = None, None

    @abc.abstractmethod
    def train(self, X, y):
        """Fit on inputs X ``(N, D)`` and targets y ``(N,)``."""

    @abc.abstractmethod
    def predict(self, X_test):
        """Predictive mean ``(N,)`` and variance ``(N,)`` at X_test ``(N, D)``."""

    def update(self, X, y):
        """Retrain from scratch on the stored data extended by the new points."""
        self.train(np.concatenate([self.X, X], axis=0), np.concatenate([self.y, y], axis=0))

    def get_json_data(self):
        as_list = lambda a: None if a is None else np.asarray(a).tolist()
        return {"X": as_list(self.X), "y": as_list(self.y), "hyperparameters": ""}

    def get_incumbent(self):
        """Best (lowest-target) observed point and its value."""
        best = int(np.argmin(self.y))
        return self.X[best], self.y[best]


def _affine(X, offset, scale):
    return (X - offset) / scale


def zero_one_normalization(X, lower=None, upper=None):
    """Map each column to [0, 1]; returns ``(X_normalized, lower, upper)``."""
    lower = np.min(X, axis=0) if lower is None else lower
    upper = np.max(X, axis=0) if upper is None else upper
    return _affine(X, lower, upper - lower), lower, upper


def zero_one_unnormalization(X_normalized, lower, upper):
    return X_normalized * (upper - lower) + lower


def zero_mean_unit_var_normalization(X, mean=None, std=None):
    """Standardise each column; returns ``(X_normalized, mean, std)``."""
    mean = np.mean(X, axis=0) if mean is None else mean
    std = np.std(X, axis=0) if std is None else std
    return _affine(X, mean, std), mean, std


def zero_mean_unit_var_unnormalization(X_normalized, mean, std):
    return X_normalized * std + mean

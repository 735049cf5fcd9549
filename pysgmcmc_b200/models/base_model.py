"""Model base class and normalisation helpers -- the parts of pysgmcmc/models/base_model.py
the BNN path uses (:66-79 shape checks, :109-137 normalisation)."""
import abc

import numpy as np


class BaseModel(object, metaclass=abc.ABCMeta):
    def __init__(self):
        self.X = None
        self.y = None

    @abc.abstractmethod
    def train(self, X, y):
        """Train the model on inputs X ``(N, D)`` and targets y ``(N,)``."""

    def update(self, X, y):
        """Retrain on the old data plus the new points (base_model.py:29-44)."""
        X = np.append(self.X, X, axis=0)
        y = np.append(self.y, y, axis=0)
        self.train(X, y)

    @abc.abstractmethod
    def predict(self, X_test):
        """Predictive mean and variance at X_test ``(N, D)``."""

    def _check_shapes_train(func):
        def func_wrapper(self, X, y, *args, **kwargs):
            assert X.shape[0] == y.shape[0]
            assert len(X.shape) == 2
            assert len(y.shape) == 1
            return func(self, X, y, *args, **kwargs)
        return func_wrapper

    def _check_shapes_predict(func):
        def func_wrapper(self, X, *args, **kwargs):
            assert len(X.shape) == 2
            return func(self, X, *args, **kwargs)
        return func_wrapper

    def get_incumbent(self):
        best_idx = np.argmin(self.y)
        return self.X[best_idx], self.y[best_idx]


def zero_one_normalization(X, lower=None, upper=None):
    if lower is None:
        lower = np.min(X, axis=0)
    if upper is None:
        upper = np.max(X, axis=0)
    return np.true_divide((X - lower), (upper - lower)), lower, upper


def zero_one_unnormalization(X_normalized, lower, upper):
    return lower + (upper - lower) * X_normalized


def zero_mean_unit_var_normalization(X, mean=None, std=None):
    if mean is None:
        mean = np.mean(X, axis=0)
    if std is None:
        std = np.std(X, axis=0)
    return (X - mean) / std, mean, std


def zero_mean_unit_var_unnormalization(X_normalized, mean, std):
    return X_normalized * std + mean

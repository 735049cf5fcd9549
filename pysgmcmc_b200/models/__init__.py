from .base_model import BaseModel
from .bayesian_neural_network import (
    BayesianNeuralNetwork,
    log_variance_prior_log_like,
    weight_prior_log_like
)

__all__ = (
    "BaseModel",
    "BayesianNeuralNetwork",
    "log_variance_prior_log_like",
    "weight_prior_log_like"
)

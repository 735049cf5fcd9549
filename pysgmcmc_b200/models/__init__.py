"""Models on top of the sampler hot path (the reference's pysgmcmc/models package): the
BOHAMIANN Bayesian neural network whose negative log likelihood and gradient are kernel K4 and
whose posterior predictive is kernel K10."""
from . import base_model as _base, bayesian_neural_network as _bnn, networks as _nets

BaseModel = _base.BaseModel
BayesianNeuralNetwork = _bnn.BayesianNeuralNetwork
log_variance_prior_log_like = _bnn.log_variance_prior_log_like
weight_prior_log_like = _bnn.weight_prior_log_like

MLPNet, TorchNet = _nets.MLPNet, _nets.TorchNet          # architectures for get_net (models/networks.py)

__all__ = ("BaseModel", "BayesianNeuralNetwork", "log_variance_prior_log_like", "weight_prior_log_like",
           "MLPNet", "TorchNet")

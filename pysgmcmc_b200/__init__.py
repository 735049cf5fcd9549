"""pysgmcmc_b200 -- a B200 (sm_100a) native SG-MCMC engine with the sampler API of
MFreidank/pysgmcmc: ``SGHMCSampler``, ``SGLDSampler``, ``RelativisticSGHMCSampler``,
``stepsize_schedules`` and the ``sample, cost = next(sampler)`` iterator, executed by
hand-written CUDA kernels behind a C ABI (include/sgmcmc_b200.h).
"""
from . import stepsize_schedules, tensor_utils, placeholders  # noqa: F401
from .placeholders import placeholder, Placeholder  # noqa: F401
from .session import Session  # noqa: F401

__version__ = "0.1.0"

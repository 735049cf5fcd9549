"""Sampler classes of the engine, under the reference's names (pysgmcmc/samplers/__init__.py):
every class keeps the reference's constructor and ``sample, cost = next(sampler)`` protocol and
runs on the CUDA kernels behind libsgmcmc_b200.so (K1-K3 element-wise updates, K6 fused target
chains, K11-K14 for the particle-coupled SVGD update)."""
from .base_classes import BurnInMCMCSampler, MCMCSampler
from .relativistic_sghmc import RelativisticSGHMCSampler
from .sghmc import SGHMCSampler
from .sgld import SGLDSampler
from .svgd import SVGDSampler

__all__ = (
    "MCMCSampler",
    "BurnInMCMCSampler",
    "SGHMCSampler",
    "SGLDSampler",
    "RelativisticSGHMCSampler",
    "SVGDSampler",
)

"""Convergence diagnostics and test densities (the reference's pysgmcmc/diagnostics package):
traces of sampler runs, effective sample sizes and Gelman-Rubin statistics computed by the GPU
reductions K8 (+ one all-reduce across GPUs), and the banana / Gaussian-mixture targets.
`multitrace` is the pymc3-free stand-in for the reference's `pymc3_multitrace`."""
from . import objective_functions  # noqa: F401
from .sample_chains import PYSGMCMCTrace, multitrace
from .sampler_diagnostics import effective_sample_sizes, gelman_rubin

__all__ = ("PYSGMCMCTrace", "multitrace", "effective_sample_sizes", "gelman_rubin", "objective_functions")

from . import objective_functions  # noqa: F401
from .sample_chains import PYSGMCMCTrace, multitrace  # noqa: F401
from .sampler_diagnostics import effective_sample_sizes, gelman_rubin  # noqa: F401

__all__ = (
    "PYSGMCMCTrace",
    "multitrace",
    "effective_sample_sizes",
    "gelman_rubin",
)

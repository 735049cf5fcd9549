from .sghmc import SGHMCSampler
from .relativistic_sghmc import RelativisticSGHMCSampler
from .sgld import SGLDSampler

__all__ = [
    "SGHMCSampler",
    "RelativisticSGHMCSampler",
    "SGLDSampler",
]

from .sghmc import SGHMCSampler
from .relativistic_sghmc import RelativisticSGHMCSampler
from .sgld import SGLDSampler
from .svgd import SVGDSampler

__all__ = [
    "SGHMCSampler",
    "RelativisticSGHMCSampler",
    "SGLDSampler",
    "SVGDSampler",
]

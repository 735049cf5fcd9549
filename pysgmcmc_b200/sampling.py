"""Sampler enumeration and factory -- same interface and error texts as
pysgmcmc/sampling.py:5-273.
"""
import inspect
from enum import Enum


class Sampler(Enum):
    """Enumeration type for all samplers we support."""

    SGHMC = "SGHMC"
    RelativisticSGHMC = "RelativisticSGHMC"
    SGLD = "SGLD"
    SVGD = "SVGD"

    @staticmethod
    def is_burn_in_mcmc(sampling_method):
        """
        >>> Sampler.is_burn_in_mcmc(Sampler.SGHMC)
        True
        >>> Sampler.is_burn_in_mcmc(Sampler.RelativisticSGHMC)
        False
        >>> Sampler.is_burn_in_mcmc(0)
        False
        >>> Sampler.is_burn_in_mcmc("test")
        False
        """
        return sampling_method in (Sampler.SGHMC, Sampler.SGLD)

    @staticmethod
    def is_supported(sampling_method):
        """Samplers the BNN model supports (sampling.py:43-64).

        >>> Sampler.is_supported(Sampler.SGHMC)
        True
        >>> Sampler.is_supported(0)
        False
        >>> Sampler.is_supported("test")
        False
        """
        return sampling_method in (Sampler.SGHMC, Sampler.SGLD)

    @classmethod
    def get_sampler(cls, sampling_method, **sampler_args):
        """Return a sampler for `sampling_method`, overriding constructor defaults with
        `sampler_args` (sampling.py:66-273).

        >>> Sampler.get_sampler(Sampler.SGHMC, unknown_argument=None, params=[], cost_fun=None)
        Traceback (most recent call last):
          ...
        ValueError: sampling.Sampler.get_sampler: 'SGHMCSampler' does not take any parameter with name 'unknown_argument' which was specified as argument to this sampler. Please ensure, that you only specify sampler arguments that fit the corresponding sampling method.
        For your choice of sampling method ('Sampler.SGHMC'), supported parameters are:
        -params
        -cost_fun
        -batch_generator
        -stepsize_schedule
        -burn_in_steps
        -mdecay
        -scale_grad
        -session
        -dtype
        -seed
        >>> Sampler.get_sampler(Sampler.SGHMC)
        Traceback (most recent call last):
          ...
        ValueError: sampling.Sampler.get_sampler: params was not overwritten as sampler argument in `sampler_args` and does not have any default value in SGHMCSampler.__init__Please pass an explicit value for this parameter.
        """
        sampler_class = _sampler_class(sampling_method)
        accepted = [name for name in inspect.signature(sampler_class.__init__).parameters
                    if name != "self"]
        defaults = {name: p.default
                    for name, p in inspect.signature(sampler_class.__init__).parameters.items()}

        for name in sampler_args:
            if name not in accepted:
                raise ValueError(_UNKNOWN_ARGUMENT.format(
                    sampler_name=sampler_class.__name__, parameter=name, sampler=sampling_method,
                    valid_parameters="\n".join("-" + a for a in accepted)))

        resolved = {}
        for name in accepted:
            if name in sampler_args:
                resolved[name] = sampler_args[name]
            elif defaults[name] is inspect.Parameter.empty:
                raise ValueError(_MISSING_ARGUMENT.format(
                    param_name=name, sampler=sampler_class.__name__))
            else:
                resolved[name] = defaults[name]
        return sampler_class(**resolved)


# exact texts of pysgmcmc/sampling.py:212-263 (pinned by the reference's doctests :136-171)
_UNKNOWN_ARGUMENT = (
    "sampling.Sampler.get_sampler: '{sampler_name}' does not take any parameter with name "
    "'{parameter}' which was specified as argument to this sampler. Please ensure, that you "
    "only specify sampler arguments that fit the corresponding sampling method.\n"
    "For your choice of sampling method ('{sampler}'), supported parameters are:\n"
    "{valid_parameters}")
_MISSING_ARGUMENT = (
    "sampling.Sampler.get_sampler: {param_name} was not overwritten as sampler argument in "
    "`sampler_args` and does not have any default value in {sampler}.__init__"
    "Please pass an explicit value for this parameter.")


def _sampler_class(sampling_method):
    from pysgmcmc_b200 import samplers
    table = {
        Sampler.SGHMC: samplers.SGHMCSampler,
        Sampler.SGLD: samplers.SGLDSampler,
        Sampler.RelativisticSGHMC: samplers.RelativisticSGHMCSampler,
        Sampler.SVGD: samplers.SVGDSampler,
    }
    if sampling_method not in table:
        raise ValueError("Sampling method {} is not supported by this engine; choose one of "
                         "{}.".format(sampling_method, ", ".join(str(k) for k in table)))
    return table[sampling_method]

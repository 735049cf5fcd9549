"""The package's doctests (the reference's main API coverage is doctests, SURVEY.md section 4:
`py.test --doctest-modules`); the ones here need no GPU."""
import doctest
import importlib

import pytest

MODULES = [
    "pysgmcmc_b200.data_batches",
    "pysgmcmc_b200.sampling",
    "pysgmcmc_b200.stepsize_schedules",
    "pysgmcmc_b200.diagnostics.objective_functions",
    "pysgmcmc_b200.samplers.base_classes",
    "pysgmcmc_b200.samplers.relativistic_sghmc",
    "pysgmcmc_b200.samplers.sghmc",
]


@pytest.mark.parametrize("name", MODULES)
def test_module_doctests(name):
    module = importlib.import_module(name)
    result = doctest.testmod(module, optionflags=doctest.ELLIPSIS | doctest.NORMALIZE_WHITESPACE)
    assert result.failed == 0, "%d doctest failure(s) in %s" % (result.failed, name)
    assert result.attempted > 0, "no doctests collected in %s" % name

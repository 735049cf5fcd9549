"""Trace adapter -- the duck-typed trace interface of
pysgmcmc/diagnostics/sample_chains.py:14-335 (`PYSGMCMCTrace`) without pymc3.
`pymc3_multitrace` (:338-384) returned a pymc3 object; here `multitrace` returns the list
of traces and the stacked device trace the GPU diagnostics consume.
"""
import logging
from itertools import islice

import numpy as np

from ..tensor_utils import get_name


class PYSGMCMCTrace(object):
    """A single chain of samples from a sampler (sample_chains.py:14-88)."""

    def __init__(self, chain_id, samples, varnames=None):
        self.chain = chain_id

        assert(hasattr(samples, "__len__")), "Samples needs to have a __len__ attribute."
        assert(len(samples) >= 1), "There needs to be at least one sample."

        self.samples = samples
        first_sample = self.samples[0]

        if isinstance(first_sample, (float, np.float32, np.float64)) or np.ndim(first_sample) == 0:
            self.n_vars = 1
            self.samples = [[sample] for sample in self.samples]
        else:
            self.n_vars = len(first_sample)

        assert(self.n_vars >= 1), "The first sample needs to have at least one variable."

        if varnames is None:
            logging.warning(
                "Variables in a trace were not named when instantiating "
                "a `pysgmcmc.diagnostics.sample_chain.PYSGMCMCTrace` "
                "from that trace. We will give them anonymous names "
                "by enumerating all target parameter dimensions."
            )
            self.varnames = [str(variable_index) for variable_index in range(self.n_vars)]
        else:
            self.varnames = varnames

        assert len(self.varnames) == self.n_vars

    @classmethod
    def from_sampler(cls, chain_id, sampler, n_samples, keep_every=1, varnames=None):
        """`n_samples` consecutive samples of `sampler` (sample_chains.py:90-176; like the
        reference, `keep_every` is accepted but not applied, :166-169)."""
        samples = [sample for sample, _ in islice(sampler, n_samples)]
        if varnames is None:
            varnames = [get_name(param, "param_%d" % i) for i, param in enumerate(sampler.params)]
        return PYSGMCMCTrace(chain_id, samples, varnames)

    def __getitem__(self, index):
        assert isinstance(index, int)
        assert 0 <= index < len(self.varnames)
        return self.get_values(self.varnames[index])

    def _slice(self, slice_):
        return PYSGMCMCTrace(chain_id=self.chain, samples=self.samples[slice_],
                             varnames=self.varnames[slice_])

    def point(self, index):
        sample = self.samples[index]
        return {varname: sample[i] for i, varname in enumerate(self.varnames)}

    def __len__(self):
        return len(self.samples)

    def get_values(self, varname, burn=0, thin=1):
        """All sampled values of variable `varname` as an ``(N, D)`` array
        (sample_chains.py:259-335)."""
        if varname not in self.varnames:
            raise ValueError(
                "Queried `PYSGMCMCTrace` for values of parameter with "
                "name '{name}' but the trace does not contain any "
                "parameter of that name. "
                "Known variable names were: '{varnames}'"
                .format(name=varname, varnames=self.varnames)
            )
        var_index = self.varnames.index(varname)
        return np.asarray([np.asarray(sample[var_index]) for sample in self.samples[burn::thin]])


def multitrace(get_sampler, n_chains=2, samples_per_chain=100, keep_every=10, parameter_names=None):
    """One trace per chain, each from a fresh ``get_sampler(session)`` like
    sample_chains.py:367-382."""
    from ..session import Session
    traces = []
    for chain_id in range(n_chains):
        sampler = get_sampler(Session())
        traces.append(PYSGMCMCTrace.from_sampler(chain_id, sampler, samples_per_chain, keep_every,
                                                 parameter_names))
    return traces

"""Relativistic SGHMC -- same constructor and iterator interface as
pysgmcmc/samplers/relativistic_sghmc.py:13-223, executed by kernel K3
(csrc/update_kernels.cu) and by K6 for the built-in test densities.
"""
import ctypes

import numpy as np
import torch

from .. import _native
from ..stepsize_schedules import ConstantStepsizeSchedule
from .base_classes import MCMCSampler


class RelativisticSGHMCSampler(MCMCSampler):
    """Relativistic stochastic-gradient HMC (Lu et al. 2017); per step and element
    (relativistic_sghmc.py:120-135, `grad` = gradient of the LOG LIKELIHOOD = -d cost)::

        vel(p) = eps * p / (m * sqrt(p*p / (m^2 c^2) + 1))
        p     += eps*grad + sqrt(eps*(2D - eps*Bhat))*N(0,1) - D*vel(p)
        theta += vel(p_new)

    Generalisation: the reference keeps one SCALAR momentum per entry of `params`
    (:108-113) and therefore only works for 1-element parameters; here the momentum is
    element-wise, which is identical for 1-element parameters.  The initial momentum is
    drawn from the relativistic marginal  p ~ exp(-m c^2 sqrt(p^2/(m^2 c^2) + 1))  which
    the reference samples with the third-party `arspy` (:143-223); see
    `_sample_relativistic_momentum`.
    """

    _STATE_NAMES = ("p",)

    def __init__(self, params, cost_fun, batch_generator=None,
                 stepsize_schedule=ConstantStepsizeSchedule(0.001),
                 mass=1.0, speed_of_light=1.0, D=1.0, Bhat=0.0,
                 session=None, dtype=torch.float32, seed=None):
        super().__init__(
            params=params, cost_fun=cost_fun, batch_generator=batch_generator,
            stepsize_schedule=stepsize_schedule,
            seed=seed, dtype=dtype, session=session
        )
        self.mass, self.speed_of_light = float(mass), float(speed_of_light)
        self.D, self.Bhat = float(D), float(Bhat)
        momentum = _sample_relativistic_momentum(
            m=self.mass, c=self.speed_of_light, n_params=self._theta.numel(), seed=self.seed)
        self._state_array("p").copy_(
            torch.as_tensor(np.asarray(momentum), dtype=self.dtype).reshape(self._theta.shape))

    @property
    def momentum(self):
        return self._views(self._state_array("p"))

    def _launch_update(self, grad, z, epsilon):
        fn = "sgmcmc_rsghmc_step_f32" if self.dtype == torch.float32 else "sgmcmc_rsghmc_step_f64"
        _native.call(fn, _native.ptr(self._theta), _native.ptr(self._state_array("p")),
                     _native.ptr(grad), _native.ptr(z), self._theta.numel(), epsilon, self.mass,
                     self.speed_of_light, self.D, self.Bhat, self._noise_seed, self.n_iterations,
                     self._elem_offset, self._stream())

    def _target_run(self, n_steps, keep_every, z, trace, costs, epsilon):
        hyper = _native.Hyper(epsilon=epsilon, mdecay=0.0, scale_grad=1.0, A=1.0, mass=self.mass,
                              speed_of_light=self.speed_of_light, D=self.D, Bhat=self.Bhat)
        _native.call("sgmcmc_target_chains_run_f32", _native.SAMPLER_RSGHMC,
                     _native.TARGET_IDS[self._native_target],
                     _native.ptr(self._theta), _native.ptr(self._state_array("p")), None, None, None,
                     None, _native.ptr(z), _native.ptr(trace), _native.ptr(costs),
                     self.n_chains, n_steps, 0, 0, keep_every, ctypes.byref(hyper),
                     self._noise_seed, self.n_iterations, self.session.chain_offset, self._stream())

    def _launch_fused_target(self, z, epsilon):
        cost = torch.empty((1, self.n_chains), dtype=self.dtype, device=self.device)
        self._target_run(1, 1, z, None, cost, epsilon)
        return cost[0] if self.multi_chain else cost[0, 0]

    def _launch_fused_run(self, n_steps, keep_every, trace, costs):
        self._target_run(n_steps, keep_every, None, trace, costs, float(next(self.stepsize_schedule)))
        self.n_iterations += n_steps


def _sample_relativistic_momentum(m, c, n_params,
                                  bounds=(float("-inf"), float("inf")),
                                  seed=None):
    """Initial values for the relativistic momentum `p` (relativistic_sghmc.py:143-223):
    `n_params` draws from the density  ~ exp(-m c^2 sqrt(p^2 / (m^2 c^2) + 1)).

    The reference uses adaptive rejection sampling from the third-party `arspy`
    (absent here; its random stream is unpinned by the reference -- only the count and
    seed determinism are tested, :189-195).  This implementation draws from the same
    density by exact rejection from a Laplace envelope: with q(p) ~ exp(-c |p|),
    K(p) = m c^2 sqrt(p^2/(m^2 c^2) + 1) >= c |p|, so accept with exp(c |p| - K(p)).

    >>> momentum_values = _sample_relativistic_momentum(m=1.0, c=1.0, n_params=10)
    >>> len(momentum_values) == 10
    True
    """
    assert isinstance(m, float)
    assert isinstance(c, float)
    lo, hi = bounds
    rng = np.random.RandomState(seed)
    out = np.empty(n_params, dtype=np.float64)
    filled = 0
    while filled < n_params:
        n = max(64, 2 * (n_params - filled))
        p = rng.laplace(0.0, 1.0 / c, size=n)
        k = m * c ** 2 * np.sqrt(p ** 2 / (m ** 2 * c ** 2) + 1.0)
        accept = (np.log(rng.uniform(size=n)) < c * np.abs(p) - k) & (p > lo) & (p < hi)
        good = p[accept][:n_params - filled]
        out[filled:filled + good.size] = good
        filled += good.size
    return list(out)

"""Adaptive SGHMC -- same constructor and iterator interface as
pysgmcmc/samplers/sghmc.py:12-251, executed by kernel K1 (csrc/update_kernels.cu),
by K6 for the built-in test densities and by K5 for the BNN cost.
"""
import ctypes

import torch

from .. import _native
from ..stepsize_schedules import ConstantStepsizeSchedule
from .base_classes import BurnInMCMCSampler


class SGHMCSampler(BurnInMCMCSampler):
    """Stochastic Gradient Hamiltonian Monte-Carlo with the burn-in adaptation of
    Springenberg et al. (2016); see pysgmcmc/samplers/sghmc.py:12-29.

    Per step and element (old = value before the step; sghmc.py:165-251)::

        r     = 1 / (tau + 1)
        tau  += safe_divide(-g*g*tau, v_hat) + 1
        minv  = safe_divide(1, safe_sqrt(v_hat))            # frozen after burn-in
        g    += -r*g + r*grad
        v_hat+= -r*v_hat + r*grad^2
        sigma = sqrt(max(2*eps_s^2*mdecay*minv - eps_s^4, 1e-16)),  eps_s = eps/sqrt(scale_grad)
        V    += -eps^2*minv*grad - mdecay*V + sigma*N(0,1)
        theta+= V
    """

    _STATE_NAMES = ("v", "tau", "g", "v_hat", "minv")

    def __init__(self, params, cost_fun, batch_generator=None,
                 stepsize_schedule=ConstantStepsizeSchedule(0.01),
                 burn_in_steps=3000, mdecay=0.05, scale_grad=1.0,
                 session=None, dtype=torch.float32, seed=None):
        super().__init__(
            params=params, cost_fun=cost_fun, burn_in_steps=burn_in_steps,
            batch_generator=batch_generator,
            seed=seed, dtype=dtype, session=session,
            stepsize_schedule=stepsize_schedule
        )
        self.mdecay = float(mdecay)
        self.scale_grad = float(scale_grad)
        # initial values: sghmc.py:126-155
        self._state_array("v").zero_()
        for name in ("tau", "g", "v_hat", "minv"):
            self._state_array(name).fill_(1.0)

    def _arrays(self):
        return [self._theta] + [self._state_array(n) for n in self._STATE_NAMES]

    def _launch_update(self, grad, z, epsilon, adapt=True):
        fn = "sgmcmc_sghmc_step_f32" if self.dtype == torch.float32 else "sgmcmc_sghmc_step_f64"
        last_burn_in = adapt and (self.burn_in_steps == 0 or self.n_iterations >= self.burn_in_steps - 1)
        store_minv = adapt and (last_burn_in or getattr(self, "track_minv", True))
        _native.call(fn, *[_native.ptr(a) for a in self._arrays()], _native.ptr(grad), _native.ptr(z),
                     self._theta.numel(), epsilon, self.mdecay, self.scale_grad,
                     int(adapt), int(store_minv), self._noise_seed, self.n_iterations,
                     self._elem_offset, self._stream())

    # ---- fused paths -------------------------------------------------------------
    def _hyper(self, epsilon):
        return _native.Hyper(epsilon=epsilon, mdecay=self.mdecay, scale_grad=self.scale_grad, A=1.0,
                             mass=1.0, speed_of_light=1.0, D=1.0, Bhat=0.0)

    def _target_run(self, n_steps, n_burn_in, keep_every, z, trace, costs, epsilon):
        hyper = self._hyper(epsilon)
        _native.call("sgmcmc_target_chains_run_f32", _native.SAMPLER_SGHMC,
                     _native.TARGET_IDS[self._native_target],
                     *[_native.ptr(a) for a in self._arrays()], _native.ptr(z),
                     _native.ptr(trace), _native.ptr(costs), self.n_chains, n_steps, n_burn_in,
                     int(self.burn_in_steps == 0), keep_every, ctypes.byref(hyper),
                     self._noise_seed, self.n_iterations, self.session.chain_offset, self._stream())

    def _launch_fused_target(self, z, epsilon, adapt=True):
        cost = torch.empty((1, self.n_chains), dtype=self.dtype, device=self.device)
        self._target_run(1, 1 if adapt else 0, 1, z, None, cost, epsilon)
        return cost[0] if self.multi_chain else cost[0, 0]

    def _launch_fused_run(self, n_steps, keep_every, trace, costs):
        self._target_run(n_steps, min(n_steps, self._burn_in_remaining()), keep_every, None, trace,
                         costs, float(next(self.stepsize_schedule)))
        self.n_iterations += n_steps

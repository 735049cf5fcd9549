"""Adaptive SGLD -- same constructor and iterator interface as
pysgmcmc/samplers/sgld.py:13-213, executed by kernel K2 (csrc/update_kernels.cu) and by
K6 for the built-in test densities.
"""
import ctypes

import torch

from .. import _native
from ..stepsize_schedules import ConstantStepsizeSchedule
from .base_classes import BurnInMCMCSampler


class SGLDSampler(BurnInMCMCSampler):
    """Stochastic Gradient Langevin Dynamics with the same burn-in preconditioner as
    SGHMC (sgld.py:149-213)::

        (r, tau, minv, g, v_hat as in SGHMC)
        sigma  = safe_sqrt(2*eps*safe_divide(minv*A, scale_grad))
        theta += -eps*minv*A*grad + sigma*N(0,1)

    Reference quirk kept on purpose: `stepsize_schedule` is accepted but NOT forwarded
    to the base class (sgld.py:96-100), so the stepsize is always the base default 0.01.
    """

    _STATE_NAMES = ("tau", "g", "v_hat", "minv")

    def __init__(self, params, cost_fun, batch_generator=None,
                 stepsize_schedule=ConstantStepsizeSchedule(0.01),
                 burn_in_steps=3000, A=1.0, scale_grad=1.0,
                 session=None, dtype=torch.float32, seed=None):
        super().__init__(
            params=params, cost_fun=cost_fun, batch_generator=batch_generator,
            burn_in_steps=burn_in_steps, seed=seed,
            session=session, dtype=dtype
        )
        self.A = float(A)
        self.scale_grad = float(scale_grad)
        for name in ("tau", "g", "v_hat", "minv"):
            self._state_array(name).fill_(1.0)

    def _arrays(self):
        return [self._theta] + [self._state_array(n) for n in self._STATE_NAMES]

    def _launch_update(self, grad, z, epsilon, adapt=True):
        fn = "sgmcmc_sgld_step_f32" if self.dtype == torch.float32 else "sgmcmc_sgld_step_f64"
        last_burn_in = adapt and (self.burn_in_steps == 0 or self.n_iterations >= self.burn_in_steps - 1)
        store_minv = adapt and (last_burn_in or getattr(self, "track_minv", True))
        _native.call(fn, *[_native.ptr(a) for a in self._arrays()], _native.ptr(grad), _native.ptr(z),
                     self._theta.numel(), epsilon, self.A, self.scale_grad,
                     int(adapt), int(store_minv), self._noise_seed, self.n_iterations,
                     self._elem_offset, self._stream())

    def _target_run(self, n_steps, n_burn_in, keep_every, z, trace, costs, epsilon):
        hyper = _native.Hyper(epsilon=epsilon, mdecay=0.0, scale_grad=self.scale_grad, A=self.A,
                              mass=1.0, speed_of_light=1.0, D=1.0, Bhat=0.0)
        theta, tau, g, v_hat, minv = self._arrays()
        _native.call("sgmcmc_target_chains_run_f32", _native.SAMPLER_SGLD,
                     _native.TARGET_IDS[self._native_target],
                     _native.ptr(theta), None, _native.ptr(tau), _native.ptr(g), _native.ptr(v_hat),
                     _native.ptr(minv), _native.ptr(z), _native.ptr(trace), _native.ptr(costs),
                     self.n_chains, n_steps, n_burn_in, int(self.burn_in_steps == 0), keep_every,
                     ctypes.byref(hyper), self._noise_seed, self.n_iterations,
                     self.session.chain_offset, self._stream())

    def _launch_fused_target(self, z, epsilon, adapt=True):
        cost = torch.empty((1, self.n_chains), dtype=self.dtype, device=self.device)
        self._target_run(1, 1 if adapt else 0, 1, z, None, cost, epsilon)
        return cost[0] if self.multi_chain else cost[0, 0]

    def _launch_fused_run(self, n_steps, keep_every, trace, costs):
        self._target_run(n_steps, min(n_steps, self._burn_in_remaining()), keep_every, None, trace,
                         costs, float(next(self.stepsize_schedule)))
        self.n_iterations += n_steps

"""Stein variational gradient descent -- same constructor and iterator interface as
pysgmcmc/samplers/svgd.py:13-182, executed by kernels K11-K14 (csrc/svgd.cu).

The particles are the engine's chain layout: every particle is one row of the flat
``[n_particles, D]`` state, and the entries of `particles` are re-pointed at those rows
(the ``tf.Variable`` behaviour).  One ``next(sampler)`` is: the cost of every particle and
its gradient (torch autograd, or a native cost such as the BNN kernel K4 with one
particle per chain), then squared distances -> exact median -> RBF kernel matrix ->
one GEMM with the Stein direction, the AdaGrad history and the update in its epilogue,
all on the sampler's stream without a host round trip.

Kept from the reference on purpose (oracle/svgd.py lists them): the gradient is that of
the COST and the step is ``-epsilon * adj_grad``; the median includes the zero diagonal;
the AdaGrad history starts at zero.
"""
import numpy as np
import torch

from .. import _native
from ..stepsize_schedules import ConstantStepsizeSchedule
from .base_classes import MCMCSampler


class SVGDSampler(MCMCSampler):
    """Stein Variational Gradient Descent Sampler (Liu & Wang 2016), svgd.py:13-148.

    particles : list of 1-d torch tensors of equal length D, one per particle.
    cost_fun : callable taking ONE particle (a ``[D]`` tensor) and returning its cost.
        A callable with ``native_cost_and_grad(theta[n, D], grad_out) -> cost[n]``
        (e.g. ``BayesianNeuralNetworkNLL``) is evaluated by its CUDA kernel instead, and a
        callable with a true ``vectorized`` attribute is called once on the ``[n, D]``
        matrix transposed to ``[D, n]`` (so ``x[0], x[1]`` index coordinates).  The built-in test
        densities (``to_negative_log_likelihood(banana_log_likelihood)``, gmm1-3) with at most
        128 particles run in the fused one-CTA kernel (``sgmcmc_svgd_target_run_f32``): one
        launch per ``next(sampler)``, ONE launch for a whole ``sampler.run(n_steps)``.
    """

    def __init__(self, particles, cost_fun, batch_generator=None,
                 stepsize_schedule=ConstantStepsizeSchedule(0.1),
                 alpha=0.9, fudge_factor=1e-6, session=None,
                 dtype=torch.float32, seed=None):
        assert isinstance(alpha, (int, float))
        assert isinstance(fudge_factor, (int, float))
        assert callable(cost_fun)
        assert dtype == torch.float32, "the SVGD kernels compute in float32"
        assert session is None or session.n_chains is None, \
            "the particles are the chains: pass them as a list, not through Session(n_chains=...)"

        particles = list(particles)
        shapes = {tuple(p.shape) for p in particles}
        if len(shapes) != 1 or len(next(iter(shapes))) != 1:
            # tf.stack(particles) must be 2-d for pdist (tensor_utils.py:390-391)
            raise ValueError('tensor_utils.pdist: A 2-d tensor must be passed.')

        self.particle_cost_fun = cost_fun

        def cost_fun_wrapper(params):
            return self._particle_costs(torch.stack(list(params)))

        cost_fun_wrapper.__name__ = getattr(cost_fun, "__name__", type(cost_fun).__name__)

        super().__init__(
            params=particles, cost_fun=cost_fun_wrapper,
            batch_generator=batch_generator,
            session=session, seed=seed, dtype=dtype,
            stepsize_schedule=stepsize_schedule
        )

        self.alpha, self.fudge_factor = float(alpha), float(fudge_factor)
        self.n_particles = n = len(particles)
        self.n_dims = D = self.n_params_per_chain // n
        #: live ``[n_particles, D]`` view of the state (svgd.py:85: tf.stack(particles))
        self.particles = self._theta.view(n, D)
        dev = self.device
        self.historical_grad = torch.zeros((n, D), dtype=torch.float32, device=dev)
        self._kernel_matrix = torch.empty((n, n), dtype=torch.float32, device=dev)
        self._kernel_sum = torch.empty((n,), dtype=torch.float32, device=dev)
        self._bandwidth = torch.zeros((4,), dtype=torch.float32, device=dev)
        # median select state + work space of the centred-Gram distance kernel
        with torch.cuda.device(dev):
            self._select_scratch = _native.svgd_scratch(n, D, dev)
        self._particles_scratch = torch.empty((n, D), dtype=torch.float32, device=dev)
        self._grad = torch.empty((n, D), dtype=torch.float32, device=dev)
        self._vmap_ok = None

    # ------------------------------------------------------------------ fused small-problem path
    def _detect_native_target(self):
        tag = getattr(self.particle_cost_fun, "native_target", None)
        if tag is None or tag[1] != -1 or not self.session.fused:
            return None
        n, D = len(self.params), self._sizes[0]
        if n > 128 or D != (2 if tag[0] == "banana" else 1):
            return None
        return tag[0]

    def _target_run(self, n_steps, keep_every, trace, costs, epsilon):
        _native.call("sgmcmc_svgd_target_run_f32", _native.TARGET_IDS[self._native_target],
                     _native.ptr(self.particles), _native.ptr(self.historical_grad), _native.ptr(trace),
                     _native.ptr(costs), self.n_particles, n_steps, keep_every, epsilon, self.alpha,
                     1. - self.alpha, self.fudge_factor, self._stream())

    def _launch_fused_target(self, z, epsilon):
        assert z is None, "SVGD is deterministic: there is no noise to inject"
        cost = torch.empty((1, self.n_particles), dtype=torch.float32, device=self.device)
        self._target_run(1, 1, None, cost, epsilon)
        return cost[0]

    # ------------------------------------------------------------------ cost + gradient
    def _particle_costs(self, X):
        """Costs of all particles, ``[n]`` (svgd.py:87-88: tf.map_fn over the particles)."""
        fun = self.particle_cost_fun
        if getattr(fun, "vectorized", False):
            return fun(X.t()).reshape(X.shape[0])
        if self._vmap_ok is not False:
            try:
                costs = torch.vmap(lambda x: torch.as_tensor(fun(x)).reshape(()))(X)
                self._vmap_ok = True
                return costs
            except Exception:
                if self._vmap_ok:           # it worked before: a genuine error of the cost function
                    raise
                self._vmap_ok = False
        return torch.stack([torch.as_tensor(fun(X[i])).reshape(()) for i in range(X.shape[0])])

    def _cost_and_grad(self):
        fun = self.particle_cost_fun
        if hasattr(fun, "native_cost_and_grad"):
            return fun.native_cost_and_grad(self.particles, self._grad), self._grad
        X = self.particles.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            costs = self._particle_costs(X)
            grad, = torch.autograd.grad(costs.sum(), X)
        self._grad.copy_(grad)
        return costs.detach(), self._grad

    # ------------------------------------------------------------------ kernels
    def _launch_kernel_matrix(self):
        _native.call("sgmcmc_svgd_kernel_matrix_f32", _native.ptr(self.particles),
                     _native.ptr(self._kernel_matrix), _native.ptr(self._kernel_sum),
                     _native.ptr(self._bandwidth), _native.ptr(self._select_scratch),
                     self._select_scratch.numel() * 8, self.n_particles, self.n_dims, self._stream())

    def svgd_kernel(self, particles=None):
        """RBF kernel matrix of the current particles and its summed gradients
        (svgd.py:150-182): returns ``(kernel_matrix [n, n], kernel_gradients [n, D])``."""
        assert particles is None or particles is self.particles, \
            "the kernel is evaluated on the sampler's own particles"
        with self._on_device():
            self._launch_kernel_matrix()
            K, X = self._kernel_matrix.clone(), self.particles
            kernel_gradients = (-(K @ X) + X * self._kernel_sum[:, None]) / self._bandwidth[2]
        return K, kernel_gradients

    @property
    def bandwidth(self):
        """``h`` of the last kernel evaluation (svgd.py:155-157), a device scalar."""
        return self._bandwidth[1]

    def _launch_update(self, grad, z, epsilon):
        assert z is None, "SVGD is deterministic: there is no noise to inject"
        self._launch_kernel_matrix()
        _native.call("sgmcmc_svgd_update_f32", _native.ptr(self.particles), _native.ptr(grad),
                     _native.ptr(self.historical_grad), _native.ptr(self._kernel_matrix),
                     _native.ptr(self._kernel_sum), _native.ptr(self._bandwidth),
                     _native.ptr(self._particles_scratch), self.n_particles, self.n_dims,
                     epsilon, self.alpha, 1. - self.alpha, self.fudge_factor, self._stream())

    # ------------------------------------------------------------------ outputs
    def _output_cost(self, cost):
        cost = cost.reshape(self.n_particles)
        return cost.detach().cpu().numpy() if self.session.output == "numpy" else cost

    def run(self, n_steps, keep_every=1):
        """`n_steps` updates without returning to the host; every `keep_every`-th particle
        set and cost vector: ``(trace [n_keep, n_particles, D], costs [n_keep, n_particles])``."""
        assert n_steps >= 0 and keep_every >= 1
        with self._on_device():
            return self._run(n_steps, keep_every)

    _prefetch_supported = False      # (the particle layout of the trace differs from the chain layout)

    def _run(self, n_steps, keep_every):
        n_keep = n_steps // keep_every
        trace = torch.empty((n_keep, self.n_particles, self.n_dims), dtype=self.dtype, device=self.device)
        costs = torch.empty((n_keep, self.n_particles), dtype=self.dtype, device=self.device)
        if n_steps > 0 and self._can_run_fused():
            self._target_run(n_steps, keep_every, trace if n_keep else None, costs if n_keep else None,
                             float(next(self.stepsize_schedule)))
            self.n_iterations += n_steps
            return trace, costs
        for s in range(n_steps):
            cost = self._step_on_device()
            if (s + 1) % keep_every == 0:
                k = (s + 1) // keep_every - 1
                trace[k].copy_(self.particles)
                costs[k].copy_(cost.reshape(self.n_particles))
        return trace, costs

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        state = super().state_dict()
        state["historical_grad"] = self.historical_grad.clone()
        return state

    def load_state_dict(self, state):
        super().load_state_dict(state)
        self.historical_grad.copy_(state["historical_grad"])

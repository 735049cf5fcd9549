"""Base classes of all samplers: same public interface as
pysgmcmc/samplers/base_classes.py (MCMCSampler :18-310, BurnInMCMCSampler :313-456),
executed by the CUDA kernels behind libsgmcmc_b200.so.

What changed underneath
-----------------------
* `params` are torch CUDA tensors.  On construction they are packed into ONE flat
  ``[C chains, D]`` buffer and every tensor in `params` is re-pointed to a view of it,
  so -- like ``tf.Variable`` objects -- they always show the sampler's current state.
* One ``next(sampler)`` = one kernel launch over all chains (two on the generic path:
  the user's cost function differentiated by torch autograd, then the fused update).
* `session` is a :class:`pysgmcmc_b200.session.Session` (device, stream, number of
  chains, output type); the remaining constructor arguments are the reference's.
* noise comes from the engine's Philox stream (seed, step, element) unless a tensor
  is fed for ``sampler.noise`` (used by the parity tests).

There is no CPU path: constructing a sampler without the CUDA library raises.
"""
import abc
import contextlib

import numpy as np
import torch

from .. import _native
from ..placeholders import Placeholder, feed
from ..session import Session
from ..stepsize_schedules import ConstantStepsizeSchedule

__all__ = (
    "MCMCSampler",
    "BurnInMCMCSampler",
)

_SUPPORTED_DTYPES = (torch.float32, torch.float64)


class MCMCSampler(object, metaclass=abc.ABCMeta):
    """Generic base class for all MCMC samplers (base_classes.py:18-310)."""

    #: names of the extra ``[C, D]`` state arrays the concrete sampler owns
    _STATE_NAMES = ()

    def __init__(self, params, cost_fun, batch_generator=None,
                 stepsize_schedule=ConstantStepsizeSchedule(0.01),
                 session=None, dtype=torch.float32, seed=None):
        # Sanitize inputs (base_classes.py:73-89)
        assert batch_generator is None or hasattr(batch_generator, "__next__")
        assert seed is None or isinstance(seed, int)
        assert session is None or isinstance(session, Session)
        assert isinstance(dtype, torch.dtype)
        assert dtype in _SUPPORTED_DTYPES, "the engine computes in float32 or float64"
        assert callable(cost_fun)
        assert hasattr(stepsize_schedule, "update")
        assert hasattr(stepsize_schedule, "__next__")
        assert hasattr(stepsize_schedule, "initial_value")

        _native.load()          # fail loudly when the CUDA library is missing

        self.dtype = dtype
        self.n_iterations = 0
        self.seed = seed
        # stream seed of the engine's Philox noise; drawn once when the user gave none
        self._noise_seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.stepsize_schedule = stepsize_schedule
        self.batch_generator = batch_generator
        self.session = Session() if session is None else session
        self.device = self.session.device
        self.params = list(params)
        self.cost_fun = cost_fun
        self.cost = None

        self._pack_params()

        self.epsilon = Placeholder("epsilon")
        self.epsilon.value = self.stepsize_schedule.initial_value
        #: feed a ``[C, D]`` (or broadcastable flat) tensor here to inject the N(0,1) draws
        self.noise = Placeholder("noise")

        self.theta_t = self.params          # live views: hold the new sample after a step
        self._native_target = self._detect_native_target()

    # ------------------------------------------------------------------ state layout
    def _pack_params(self):
        C = self.session.n_chains
        self.multi_chain = C is not None
        self.n_chains = C if self.multi_chain else 1
        C = self.n_chains
        assert len(self.params) > 0
        shapes, sizes = [], []
        for p in self.params:
            assert isinstance(p, torch.Tensor), "params must be torch tensors"
            if self.multi_chain:
                assert p.dim() >= 1 and p.shape[0] == C, \
                    "with Session(n_chains=C) every parameter needs a leading axis of size C"
                shapes.append(tuple(p.shape[1:]))
            else:
                shapes.append(tuple(p.shape))
            sizes.append(int(np.prod(shapes[-1], dtype=np.int64)))
        self._shapes, self._sizes = shapes, sizes
        self._offsets = [int(o) for o in np.cumsum([0] + sizes[:-1])]
        self.n_params_per_chain = D = int(sum(sizes))
        self._theta = torch.empty((C, D), dtype=self.dtype, device=self.device)
        for p, off, n in zip(self.params, self._offsets, sizes):
            self._theta[:, off:off + n] = p.detach().to(device=self.device, dtype=self.dtype).reshape(C, n)
        # re-point the user's tensors at the flat state (the tf.Variable behaviour)
        for p, view in zip(self.params, self._views(self._theta)):
            p.data = view
        self.vectorized_params = self._views(self._theta, vectorized=True)
        if self._STATE_NAMES:
            self._state = torch.empty((len(self._STATE_NAMES), C, D), dtype=self.dtype, device=self.device)
        self._grad = None

    def _views(self, flat, vectorized=False):
        """Per-parameter views of a flat ``[C, D]`` buffer, in the user's shapes
        (or ``(n, 1)`` like tensor_utils.vectorize when `vectorized`)."""
        # one split instead of a slice per parameter: this runs on every `next()`
        if self.multi_chain:
            pieces = flat.split(self._sizes, dim=1)
            return [v.view((self.n_chains,) + ((n, 1) if vectorized else shp))
                    for v, shp, n in zip(pieces, self._shapes, self._sizes)]
        pieces = flat[0].split(self._sizes)
        return [v.view((n, 1) if vectorized else shp) for v, shp, n in zip(pieces, self._shapes, self._sizes)]

    def _state_array(self, name):
        return self._state[self._STATE_NAMES.index(name)]

    # ------------------------------------------------------------------ cost + gradient
    def _detect_native_target(self):
        tag = getattr(self.cost_fun, "native_target", None)
        if tag is None or tag[1] != -1 or self.dtype != torch.float32 or not self.session.fused:
            return None
        name = tag[0]
        n_scalars = 2 if name == "banana" else 1       # the targets take scalar parameters
        ok = len(self._shapes) == n_scalars and all(s in ((), (1,)) for s in self._shapes)
        return name if ok else None

    def _autograd_cost_and_grad(self):
        """Generic path: differentiate the user's callable with torch autograd and pack
        the gradients into the flat ``[C, D]`` layout the update kernel reads."""
        leaves = [p.detach().requires_grad_(True) for p in self.params]
        with torch.enable_grad():
            cost = self.cost_fun(leaves)
            if not isinstance(cost, torch.Tensor):
                raise TypeError("cost_fun must return a torch tensor")
            if self.multi_chain:
                cost = cost.reshape(self.n_chains, -1).sum(dim=1)
            else:
                cost = cost.reshape(())
            grads = torch.autograd.grad(cost.sum(), leaves, allow_unused=True)
        if self._grad is None:
            self._grad = torch.empty_like(self._theta)
        for gview, g in zip(self._views(self._grad), grads):
            if g is None:
                gview.zero_()
            else:
                gview.copy_(g.reshape(gview.shape))
        return cost.detach(), self._grad

    def _native_cost_and_grad(self):
        """Cost functions implemented in CUDA expose
        ``native_cost_and_grad(theta_flat, grad_out) -> cost [C]``."""
        if self._grad is None:
            self._grad = torch.empty_like(self._theta)
        cost = self.cost_fun.native_cost_and_grad(self._theta, self._grad)
        return cost, self._grad

    def _cost_and_grad(self):
        if (hasattr(self.cost_fun, "native_cost_and_grad") and self.dtype == torch.float32
                and getattr(self.cost_fun, "supports_native", True)):
            return self._native_cost_and_grad()
        return self._autograd_cost_and_grad()

    # ------------------------------------------------------------------ per-step inputs
    def _next_batch(self):
        """Next minibatch feed (base_classes.py:124-192); ``{}`` without a generator.

        >>> MCMCSampler._next_batch(type("S", (), {"batch_generator": None})())
        {}
        """
        if self.batch_generator is not None:
            return next(self.batch_generator)
        return dict()

    def _next_stepsize(self):
        epsilon = next(self.stepsize_schedule)
        return {self.epsilon: epsilon}

    def _noise_tensor(self):
        """Injected noise as a flat ``[C, D]`` tensor, or None for the in-kernel Philox stream."""
        z = self.noise.value
        if z is None:
            return None
        self.noise.value = None     # an injected draw is consumed by exactly one step
        if isinstance(z, (list, tuple)):
            z = torch.cat([torch.as_tensor(zi).reshape(self.n_chains, -1) for zi in z], dim=1)
        z = torch.as_tensor(z).to(device=self.device, dtype=self.dtype)
        return z.reshape(self.n_chains, self.n_params_per_chain).contiguous()

    @property
    def _elem_offset(self):
        return self.session.chain_offset * self.n_params_per_chain

    def _stream(self):
        return _native.stream_ptr(self.session.stream, self._device_index())

    def _device_index(self):
        idx = self.__dict__.get("_dev_idx")
        if idx is None:
            d = torch.device(self.device)
            idx = self.__dict__["_dev_idx"] = d.index if d.index is not None else torch.cuda.current_device()
        return idx

    @contextlib.contextmanager
    def _on_device(self):
        """Everything a step does -- the cost / gradient (autograd or K4), the minibatch index
        kernel, the update kernel, the snapshot of the sample -- runs on the session's device
        and, when ``Session(stream=s)`` names one, on THAT stream (torch ops follow the current
        stream, the native calls get its handle), so the producers and consumers of `grad` and
        `theta` are ordered on one stream."""
        stream = self.session.stream
        if stream is None and torch.cuda.current_device() == self._device_index():
            yield                      # already there (the nested uses inside one step): nothing to switch
            return
        with torch.cuda.device(self.device):
            if stream is not None:
                if not getattr(self, "_stream_joined", False):
                    # the constructor(s) filled the state on the stream that was current then
                    stream.wait_stream(torch.cuda.current_stream(self.device))
                    self._stream_joined = True
                with torch.cuda.stream(stream):
                    yield
            else:
                yield

    # ------------------------------------------------------------------ outputs
    def _output_params(self):
        if self.session.output == "numpy":
            host = self._theta.detach().cpu().numpy()
            outs = []
            for shp, off, n in zip(self._shapes, self._offsets, self._sizes):
                v = host[:, off:off + n]
                outs.append(v.reshape((self.n_chains,) + shp).copy() if self.multi_chain
                            else v[0].reshape(shp).copy())
            return outs
        snapshot = self._theta.clone()
        return self._views(snapshot)

    def _output_cost(self, cost):
        if self.session.output == "numpy":
            c = cost.detach().cpu().numpy()
            return c if self.multi_chain else c.reshape(())[()]
        return cost if self.multi_chain else cost.reshape(())

    # ------------------------------------------------------------------ iterator protocol
    def __iter__(self):
        return self

    @abc.abstractmethod
    def _launch_update(self, grad, z, epsilon, **kwargs):
        """Launch the sampler's update kernel on the flat state."""

    def _launch_fused_target(self, z, epsilon, **kwargs):
        raise NotImplementedError

    def _advance(self, feed_dict, **kwargs):
        """One step on the device: feed, cost/gradient at the OLD theta, update kernel.
        Returns the cost at the pre-update point (base_classes.py:298-300)."""
        feed(feed_dict)
        epsilon = float(self.epsilon.value)
        with self._on_device():
            z = self._noise_tensor()
            if self._native_target is not None:
                cost = self._launch_fused_target(z, epsilon, **kwargs)
            else:
                cost, grad = self._cost_and_grad()
                self._launch_update(grad, z, epsilon, **kwargs)
        self.cost = cost
        return cost

    def __next__(self, feed_dict=None):
        """Compute and return the next sample and the cost of the previous one
        (base_classes.py:258-310): ``sample, cost = next(sampler)``."""
        assert (feed_dict is None or hasattr(feed_dict, "update"))
        assert hasattr(self, "theta_t") or not hasattr(self, "cost")

        if self._prefetching(feed_dict):
            return self._next_prefetched(unwrap=True)
        self._drop_prefetched()

        if feed_dict is None:
            feed_dict = dict()

        with self._on_device():
            feed_dict.update(self._next_batch())
            feed_dict.update(self._next_stepsize())
            cost = self._advance(feed_dict)
            params, cost = self._output_params(), self._output_cost(cost)

        if len(params) == 1:
            # unravel single-element lists to scalars (base_classes.py:302-304)
            params = params[0]

        self.stepsize_schedule.update(params, cost)

        self.n_iterations += 1

        return params, cost

    # ------------------------------------------------------------------ next() from prefetched steps
    _pf = None
    _prefetch_supported = True

    def _prefetching(self, feed_dict):
        return (not feed_dict and self._prefetch_supported and self.session.prefetch > 1
                and self._can_run_fused())

    def _next_prefetched(self, unwrap):
        """Session(prefetch=S): the next (sample, cost) pair out of a block of S steps computed by
        one launch of the fused kernels (`run`).  `n_iterations` counts the pairs handed out, the
        device state is `_pf["ahead"]` steps further."""
        pf = self._pf
        if pf is None or pf["pos"] == pf["n"]:
            S = self.session.prefetch
            if self._burn_in_remaining_for_prefetch() > 0:
                # a block never crosses the end of burn-in: `next()` changes its return type there
                S = min(S, self._burn_in_remaining_for_prefetch())
            handed_out = self.n_iterations
            with self._on_device():
                trace, costs = self._run(S, 1)           # advances n_iterations by S
                if self.session.output == "numpy":
                    trace, costs = trace.cpu().numpy(), costs.cpu().numpy()
            self.n_iterations = handed_out
            pf = self._pf = {"trace": trace, "costs": costs, "pos": 0, "n": S}
        i = pf["pos"]
        pf["pos"] += 1
        flat = pf["trace"][i]
        if self.session.output == "numpy":
            params = []
            for shp, off, n in zip(self._shapes, self._offsets, self._sizes):
                v = flat[:, off:off + n]
                params.append(v.reshape((self.n_chains,) + shp) if self.multi_chain else v[0].reshape(shp))
            c = pf["costs"][i]
            cost = c if self.multi_chain else c.reshape(())[()]
        else:
            params = self._views(flat)
            cost = self._output_cost(pf["costs"][i])
        if unwrap and len(params) == 1:
            params = params[0]
        self.stepsize_schedule.update(params, cost)
        self.n_iterations += 1
        return params, cost

    def _burn_in_remaining_for_prefetch(self):
        return 0

    def _drop_prefetched(self):
        """Leaving the prefetched mode (something was fed, `run()` was called, ...): the steps that
        were computed ahead but not handed out yet are skipped -- the chain continues from the
        device state."""
        pf = self._pf
        if pf is not None:
            self.n_iterations += pf["n"] - pf["pos"]
            self._pf = None

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        self._drop_prefetched()
        return self._state_dict()

    def _state_dict(self):
        """Everything needed to continue this chain bit-identically: the flat state arrays,
        the iteration counter (= Philox step), the noise seed and, for on-device minibatch
        generators, the MT19937 streams.  (The reference keeps its state in TF session
        variables and has no checkpointing)"""
        state = {
            "theta": self._theta.clone(),
            "n_iterations": self.n_iterations,
            "noise_seed": self._noise_seed,
            "epsilon": self.epsilon.value,
        }
        if self._STATE_NAMES:
            state["state"] = self._state.clone()
            state["state_names"] = tuple(self._STATE_NAMES)
        gen = self.batch_generator
        if gen is not None and hasattr(gen, "state_dict"):
            state["batch_generator"] = gen.state_dict()
        return state

    def load_state_dict(self, state):
        assert tuple(state["theta"].shape) == tuple(self._theta.shape), "layout mismatch"
        self._pf = None
        self._theta.copy_(state["theta"])
        if self._STATE_NAMES:
            assert tuple(state["state_names"]) == tuple(self._STATE_NAMES)
            self._state.copy_(state["state"])
        self.n_iterations = int(state["n_iterations"])
        self._noise_seed = int(state["noise_seed"])
        self.epsilon.value = state["epsilon"]
        gen = self.batch_generator
        if "batch_generator" in state and gen is not None and hasattr(gen, "load_state_dict"):
            gen.load_state_dict(state["batch_generator"])

    # ------------------------------------------------------------------ device-resident runs
    def _can_run_fused(self):
        return (self._native_target is not None and self.batch_generator is None
                and type(self.stepsize_schedule) is ConstantStepsizeSchedule)

    def run(self, n_steps, keep_every=1):
        """Advance all chains `n_steps` steps WITHOUT returning to the host in between.

        Returns ``(trace, costs)``: device tensors ``[n_steps // keep_every, C, D]`` and
        ``[n_steps // keep_every, C]`` holding every `keep_every`-th (sample, cost) pair
        -- the pairs ``next(sampler)`` would have returned.  With a built-in target and a
        constant stepsize this is ONE kernel launch (K6); otherwise it loops the per-step
        kernels on the device.
        """
        assert n_steps >= 0 and keep_every >= 1
        self._drop_prefetched()
        with self._on_device():
            return self._run(n_steps, keep_every)

    def _run(self, n_steps, keep_every):
        n_keep = n_steps // keep_every
        C, D = self.n_chains, self.n_params_per_chain
        trace = torch.empty((n_keep, C, D), dtype=self.dtype, device=self.device)
        costs = torch.empty((n_keep, C), dtype=self.dtype, device=self.device)
        if n_steps == 0:
            return trace, costs
        if self._can_run_fused():
            self._launch_fused_run(n_steps, keep_every, trace, costs)
            return trace, costs
        for s in range(n_steps):
            cost = self._step_on_device()
            if (s + 1) % keep_every == 0:
                k = (s + 1) // keep_every - 1
                trace[k].copy_(self._theta)
                costs[k].copy_(cost.reshape(C))
        return trace, costs

    def _step_on_device(self):
        feed_dict = dict()
        feed_dict.update(self._next_batch())
        feed_dict.update(self._next_stepsize())
        cost = self._advance(feed_dict)
        self.stepsize_schedule.update(self.params, cost)
        self.n_iterations += 1
        return cost

    def _launch_fused_run(self, n_steps, keep_every, trace, costs):
        raise NotImplementedError


class BurnInMCMCSampler(MCMCSampler, metaclass=abc.ABCMeta):
    """Base class for samplers that adapt their mass matrix during a burn-in phase
    (base_classes.py:313-456)."""

    def __init__(self, params, cost_fun, batch_generator=None,
                 stepsize_schedule=ConstantStepsizeSchedule(0.01),
                 burn_in_steps=3000,
                 session=None, dtype=torch.float32, seed=None):
        # Sanitize inputs
        assert isinstance(burn_in_steps, int)

        super().__init__(params=params, cost_fun=cost_fun,
                         stepsize_schedule=stepsize_schedule,
                         batch_generator=batch_generator,
                         seed=seed, dtype=dtype, session=session)

        self.burn_in_steps = burn_in_steps

    @property
    def is_burning_in(self) -> bool:
        """`True` while ``n_iterations < burn_in_steps`` (base_classes.py:393-406)."""
        return self.n_iterations < self.burn_in_steps

    @property
    def _adapts(self):
        # burn_in_steps == 0 never freezes the mass matrix (base_classes.py:449)
        return self.is_burning_in or self.burn_in_steps == 0

    @property
    def minv_t(self):
        """Per-parameter views of the inverse mass matrix used by the last step."""
        return self._views(self._state_array("minv"))

    @property
    def minv(self):
        """The mass matrix inverse fetched by the last burn-in step (base_classes.py:438);
        frozen and re-used once burn-in is over (:448-454)."""
        views = self.minv_t
        if self.session.output == "numpy":
            return [v.detach().cpu().numpy().copy() for v in views]
        return [v.clone() for v in views]

    def __next__(self, feed_dict=None):
        """One sampler step (base_classes.py:408-456).  During burn-in the mass matrix is
        adapted and the sample is always returned as a list; afterwards the frozen `minv`
        is used, a caller-supplied `feed_dict` is dropped (:454) and single-parameter
        samples are unwrapped (:302-304)."""
        assert (feed_dict is None or hasattr(feed_dict, "update"))

        if self._prefetching(feed_dict):
            # (burn-in steps return the parameter list as it is, later ones unwrap: :446 / :302-304)
            return self._next_prefetched(unwrap=not self.is_burning_in)
        self._drop_prefetched()

        if feed_dict is None:
            feed_dict = dict()

        if self.is_burning_in:
            with self._on_device():
                feed_dict.update(self._next_batch())
                feed_dict.update(self._next_stepsize())

                cost = self._advance(feed_dict, adapt=True)
                params, cost = self._output_params(), self._output_cost(cost)

            self.stepsize_schedule.update(params, cost)

            self.n_iterations += 1
            return params, cost

        if self.burn_in_steps > 0:
            # the reference REPLACES the caller's feed with the frozen minv here
            noise = feed_dict.get(self.noise) if hasattr(feed_dict, "get") else None
            feed_dict = dict() if noise is None else {self.noise: noise}

        return super().__next__(feed_dict=feed_dict)

    def _advance(self, feed_dict, adapt=None, **kwargs):
        if adapt is None:
            adapt = self._adapts
        return super()._advance(feed_dict, adapt=adapt, **kwargs)

    def _burn_in_remaining(self):
        return max(0, self.burn_in_steps - self.n_iterations)

    def _burn_in_remaining_for_prefetch(self):
        return self._burn_in_remaining()

"""Adaptive SGHMC -- same constructor and iterator interface as
pysgmcmc/samplers/sghmc.py:12-251, executed by kernel K1 (csrc/update_kernels.cu),
by K6 for the built-in test densities and by K5 for the BNN cost.
"""
import ctypes

import torch

from .. import _native
from ..stepsize_schedules import ConstantStepsizeSchedule
from . import base_classes
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
        # theta and the state rows are allocated once (load_state_dict copies into them): built once, this
        # list is on the per-step path of `next()`
        arrays = self.__dict__.get("_arrays_cache")
        if arrays is None or arrays[0] is not self._theta:
            arrays = self.__dict__["_arrays_cache"] = [self._theta] + [self._state_array(n) for n in self._STATE_NAMES]
        return arrays

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

    def _bnn_run_ok(self):
        """The BNN cost with device-resident data: K4 + K1 can be driven from C (K5)."""
        from ..data_batches import DeviceBatchGenerator
        cf = self.cost_fun
        return (getattr(cf, "bnn_native", False) and self.dtype == torch.float32 and self.session.fused
                and getattr(cf, "X", None) is not None
                and type(self.stepsize_schedule) is ConstantStepsizeSchedule
                and (self.batch_generator is None or isinstance(self.batch_generator, DeviceBatchGenerator)))

    def _can_run_fused(self):
        return super()._can_run_fused() or self._bnn_run_ok()

    #: Few chains: every chain lives on one SM (csrc/bnn_resident.cu) -- the whole state in shared memory for a
    #: block of `run()` steps, one launch per `next()` / `iter_host` step -- instead of K4 then K1 streaming all
    #: chains through HBM every step.  Used when the sampler has at most this many chains (None: 2 per SM of
    #: the device -- beyond that a one-step call, which loads and stores the whole state, loses to K4 + K1; 0: never) and the shape fits (float32, get_default_net, minibatch <= 32).  One
    #: sampler uses ONE of the two arithmetics everywhere, so run(n) == n x next() == iter_host bit for bit
    #: either way; the two differ from each other in the rounding of the gradient's dot products only
    #: (both within 1e-5 of the oracle after 1000 steps, tests/test_bnn_resident_gpu.py, tests/test_bnn_gpu.py).
    RESIDENT_MAX_CHAINS = None

    _sm_count = {}

    def _resident_limit(self):
        if self.RESIDENT_MAX_CHAINS is not None:
            return self.RESIDENT_MAX_CHAINS
        key = str(self.device)
        if key not in SGHMCSampler._sm_count:          # (a device query costs tens of microseconds: once)
            SGHMCSampler._sm_count[key] = torch.cuda.get_device_properties(self.device).multi_processor_count
        return 2 * SGHMCSampler._sm_count[key]

    def _resident_ok(self, batch):
        cf = self.cost_fun
        if not (getattr(cf, "bnn_native", False) and self.dtype == torch.float32 and self.session.fused
                and 0 < self.n_chains <= self._resident_limit()):
            return False
        fits = self.__dict__.setdefault("_resident_fits", {})
        key = (int(cf.n_in), int(batch))
        if key not in fits:
            fits[key] = _native.load().sgmcmc_bnn_resident_supported(*key) == 1
        return fits[key]

    def _advance(self, feed_dict, adapt=None, **kwargs):
        """One `next()` step; with few chains the whole step is one launch of the resident kernel."""
        cf = self.cost_fun
        if not (getattr(cf, "bnn_native", False) and self.dtype == torch.float32 and self._native_target is None
                and 0 < self.n_chains <= self._resident_limit() and self.session.fused):
            return super()._advance(feed_dict, adapt=adapt, **kwargs)
        base_classes.feed(feed_dict)
        if adapt is None:
            adapt = self._adapts
        with self._on_device():
            X, y, starts, batch = cf._device_batch()
            if not self._resident_ok(batch):
                return super()._advance({}, adapt=adapt, **kwargs)
            epsilon = float(self.epsilon.value)
            z = self._noise_tensor()
            if self._grad is None:
                self._grad = torch.empty_like(self._theta)
            cost = torch.empty(self.n_chains, dtype=self.dtype, device=self.device)
            _native.call("sgmcmc_bnn_sghmc_run_resident_f32", *[_native.ptr(a) for a in self._arrays()],
                         _native.ptr(X), _native.ptr(y), _native.ptr(starts), _native.ptr(z), None, None, None,
                         _native.ptr(cost), _native.ptr(self._grad), self.n_chains, cf.n_in, batch,
                         float(cf.batch_size), cf.n_examples, 1, 1 if adapt else 0, 0, 1, epsilon, self.mdecay,
                         self.scale_grad, self._noise_seed, self.n_iterations, self.session.chain_offset,
                         self._stream())
        self.cost = cost
        return cost

    def _launch_fused_run(self, n_steps, keep_every, trace, costs):
        epsilon = float(next(self.stepsize_schedule))
        n_burn_in = min(n_steps, self._burn_in_remaining())
        if self._native_target is not None:
            self._target_run(n_steps, n_burn_in, keep_every, None, trace, costs, epsilon)
            self.n_iterations += n_steps
        else:
            self._bnn_run(n_steps, n_burn_in, keep_every, trace, costs, epsilon)

    #: steps per K5 call when the minibatch indices come from the device generator: K7 for
    #: the next chunk runs on a side stream while the main stream samples the current one
    RUN_CHUNK = 128

    def _bnn_run(self, n_steps, n_burn_in, keep_every, trace, costs, epsilon):
        """K5 (csrc/bnn.cu: sgmcmc_bnn_sghmc_run_f32) in chunks of whole thinning periods, with
        the on-device minibatch indices (K7) of chunk i+1 generated concurrently with chunk i."""
        cf, gen = self.cost_fun, self.batch_generator
        # no index generator: every step evaluates the whole resident dataset (what the
        # per-step path and the differentiable cost do); raises when it is too large
        batch = cf.actual_batch if gen is not None else cf.full_dataset_batch()
        resident = self._resident_ok(batch)
        if self._grad is None and not resident:
            self._grad = torch.empty_like(self._theta)
        cost_scratch = torch.empty(self.n_chains, dtype=self.dtype, device=self.device)
        if gen is None or n_steps <= self.RUN_CHUNK:
            chunk = n_steps
        elif n_steps < keep_every:
            chunk = self.RUN_CHUNK
        else:
            chunk = keep_every * max(1, self.RUN_CHUNK // keep_every)
        main = self.session.stream if self.session.stream is not None else torch.cuda.current_stream(self.device)
        C, D = self.n_chains, self.n_params_per_chain
        pending = None if gen is None else gen.next_block_async(min(chunk, n_steps))
        done = 0
        while done < n_steps:
            n = min(chunk, n_steps - done)
            starts = None
            if gen is not None:
                starts, ready = pending
                main.wait_event(ready)
                starts.record_stream(main)
            k0 = done // keep_every                       # chunks start on a thinning boundary
            tr = trace[k0:] if trace is not None and k0 < trace.shape[0] else None
            co = costs[k0:] if costs is not None and k0 < costs.shape[0] else None
            # every chain resident on an SM for the whole chunk (csrc/bnn_resident.cu), or K4 then K1 per step
            fn, work = (("sgmcmc_bnn_sghmc_run_resident_f32", (None, _native.ptr(cost_scratch), None)) if resident else
                        ("sgmcmc_bnn_sghmc_run_f32", (_native.ptr(self._grad), _native.ptr(cost_scratch))))
            _native.call(fn, *[_native.ptr(a) for a in self._arrays()],
                         _native.ptr(cf.X), _native.ptr(cf.y), _native.ptr(starts), None,
                         _native.ptr(tr), _native.ptr(co), *work, C, cf.n_in, batch, float(cf.batch_size),
                         cf.n_examples, n, min(n, max(0, n_burn_in - done)), int(self.burn_in_steps == 0),
                         keep_every, epsilon, self.mdecay, self.scale_grad, self._noise_seed,
                         self.n_iterations, self.session.chain_offset, self._stream())
            if gen is not None:
                # the indices of the next chunk, on the side stream -- requested AFTER this chunk's launch so that
                # its CTAs get their SMs first (a resident CTA fills an SM: with one chain per SM an index
                # kernel that got there earlier would push chains into a second wave)
                pending = gen.next_block_async(min(chunk, n_steps - done - n)) if done + n < n_steps else None
            self.n_iterations += n
            done += n
        self.cost = cost_scratch

    # ---- next(sampler) with HOST minibatches, pipelined ---------------------------------
    def iter_host(self, host_starts, sample_every=None, lookahead=4, sample_phase=0):
        """Generator over ``(sample, cost)`` like ``next(sampler)``, for the BNN cost with the
        dataset resident on the device and the minibatch choice made on the HOST: row s of
        `host_starts` (pinned int32 ``[n_steps, C]``) holds the start index of every chain's
        minibatch of step s (``x[start:start+B]``, data_batches.py:120-123).

        Every step copies its row host->device, runs K4 + K1 and copies the per-chain cost
        back into pinned host memory; every `sample_every`-th step the whole sample ``[C, D]``
        comes back as well (``None`` otherwise) -- the thinning of
        ``BayesianNeuralNetwork.train`` (bayesian_neural_network.py:510-531); `sample_phase` shifts
        which steps those are (a call that continues a thinning period started earlier: the samples
        are the steps s with ``(s + 1 + sample_phase) % sample_every == 0``).  Unlike a plain
        ``next()`` loop the device never waits for the host: up to `lookahead` further steps
        are already queued when ``(sample_s, cost_s)`` is yielded, the copies run on their own
        streams (one per direction) and the sample is snapshotted device-to-device before it
        travels, so the result is the same as the synchronous loop, bit for bit.  The pipeline
        itself is native (csrc/host_pipeline.cu: two ctypes calls per step).

        The yielded arrays are views of pinned buffers: a cost row is re-used `lookahead + 1`
        yields later, a sample buffer only after the NEXT sample was yielded (every sample that
        can be in flight while the caller still holds one has its own pinned slot) -- copy what
        must be kept longer.
        """
        assert self._bnn_run_ok() and self._native_target is None, "iter_host needs the native BNN cost"
        assert host_starts.dtype == torch.int32 and host_starts.dim() == 2 and host_starts.shape[1] == self.n_chains
        assert host_starts.is_pinned() and host_starts.is_contiguous(), "host_starts must be pinned host memory"
        assert 0 <= lookahead < 64
        assert sample_every is None or sample_every >= 1
        cf, C, D = self.cost_fun, self.n_chains, self.n_params_per_chain
        n_steps, depth = host_starts.shape[0], lookahead + 1
        if self.RESIDENT_HOST_BLOCKS and self._resident_ok(cf.actual_batch):
            yield from self._iter_host_blocks(host_starts, sample_every, lookahead, sample_phase)
            return
        if self._grad is None:
            with self._on_device():
                self._grad = torch.empty_like(self._theta)
        # samples that can be queued before the caller lets go of the one it was handed: the
        # caller holds sample k (step s_k) until it asks for step s_k + 1, by which time steps up
        # to s_k + depth are enqueued; slot k is re-used by sample k + n_slots, at step
        # s_k + n_slots * sample_every >= s_k + depth + sample_every
        n_slots = min(depth + 1, depth // sample_every + 2) if sample_every else 0
        handle, h_cost, h_sample = self._host_pipeline(depth, n_slots, self._resident_ok(cf.actual_batch))
        epsilon = float(next(self.stepsize_schedule))
        arrays = [_native.ptr(a) for a in self._arrays()]
        X, y, grad = _native.ptr(cf.X), _native.ptr(cf.y), _native.ptr(self._grad)
        starts0, row_bytes = host_starts.data_ptr(), C * 4
        cost0 = h_cost.data_ptr()
        sample0, sample_bytes = (h_sample.data_ptr(), C * D * 4) if sample_every else (None, 0)
        ticket = ctypes.c_int64()
        sample_slot = [-1] * depth
        first = None

        def enqueue(s):
            nonlocal first
            b = s % depth
            wants = bool(sample_every) and (s + 1 + sample_phase) % sample_every == 0
            # samples are numbered over the life of the sampler, so a second iter_host() call
            # does not start on the slot the caller may still be reading
            sample_slot[b] = (self._host_samples % n_slots) if wants else -1
            self._host_samples += int(wants)
            with self._on_device():          # per call, not across yields: the caller's device stays its own
                _native.call("sgmcmc_bnn_host_pipeline_step", handle, *arrays, X, y, starts0 + s * row_bytes,
                             cost0 + b * row_bytes, sample0 + sample_slot[b] * sample_bytes if wants else None,
                             grad, cf.n_in, cf.actual_batch, float(cf.batch_size), cf.n_examples,
                             min(2, max(0, self.burn_in_steps - self.n_iterations)),
                             int(self.burn_in_steps == 0), epsilon, self.mdecay, self.scale_grad,
                             self._noise_seed, self.n_iterations, self.session.chain_offset, self._stream(),
                             ctypes.byref(ticket))
            if first is None:
                first = ticket.value                  # tickets count over the life of the handle
            self.n_iterations += 1

        queued = 0
        for s in range(n_steps):
            while queued < n_steps and queued <= s + lookahead:
                enqueue(queued)
                queued += 1
            with torch.cuda.device(self.device):
                _native.call("sgmcmc_bnn_host_pipeline_wait", handle, first + s)   # step s is in host memory
            b = s % depth
            yield (h_sample[sample_slot[b]].numpy() if sample_slot[b] >= 0 else None), h_cost[b].numpy()

    _host_samples = 0

    #: `iter_host` of a sampler on the resident kernel runs BLOCKS of up to lookahead + 1 steps per launch (a
    #: block ends at a sample step) instead of one launch per step; False: one launch per step.
    RESIDENT_HOST_BLOCKS = True

    @staticmethod
    def _host_blocks(n_steps, max_steps, sample_every=None, sample_phase=0):
        """Cut steps 0 .. n_steps - 1 into blocks ``[s0, s1)`` of at most `max_steps` steps such that every sample
        step (``(s + 1 + sample_phase) % sample_every == 0``) is the LAST step of its block: the sample is the
        state a block's launch leaves behind.

        >>> SGHMCSampler._host_blocks(10, 4, sample_every=6, sample_phase=1)
        [(0, 4), (4, 5), (5, 9), (9, 10)]
        >>> SGHMCSampler._host_blocks(5, 2)
        [(0, 2), (2, 4), (4, 5)]
        """
        bounds, s0 = [], 0
        while s0 < n_steps:
            s1 = min(s0 + max_steps, n_steps)
            if sample_every:
                first_sample = s0 + (-(s0 + 1 + sample_phase)) % sample_every      # first sample step >= s0
                s1 = min(s1, first_sample + 1)
            bounds.append((s0, s1))
            s0 = s1
        return bounds

    def _iter_host_blocks(self, host_starts, sample_every, lookahead, sample_phase):
        """`iter_host` for few chains: the chains stay on their SMs (csrc/bnn_resident.cu) for a block of up to
        lookahead + 1 steps -- one host-to-device copy of the block's index rows, one launch, one device-to-host
        copy of its cost rows (and of the sample a block ends with) -- two blocks queued ahead of the one being
        handed out.  Same pairs as stepping one by one, bit for bit; the buffer lifetimes promised by
        `iter_host` hold (a cost row lives for lookahead + 1 further yields, a sample until the next one)."""
        cf, C, D, dev = self.cost_fun, self.n_chains, self.n_params_per_chain, self.device
        n_steps, S = host_starts.shape[0], lookahead + 1
        bounds = self._host_blocks(n_steps, S, sample_every, sample_phase)
        shortest = min(S, sample_every) if sample_every else S
        n_slots = -(-(S + 1) // shortest) + 3            # blocks a cost row has to outlive, + the two queued ahead
        n_sample_slots = 4 if sample_every else 0
        key = ("blocks", S, n_slots, n_sample_slots)
        pipe = getattr(self, "_host_block_pipe", None)
        if pipe is None or pipe["key"] != key:
            with torch.cuda.device(dev):
                pipe = self._host_block_pipe = {
                    "key": key, "s_in": torch.cuda.Stream(device=dev), "s_out": torch.cuda.Stream(device=dev),
                    "d_starts": torch.empty((n_slots, S, C), dtype=torch.int32, device=dev),
                    "d_cost": torch.empty((n_slots, S, C), dtype=self.dtype, device=dev),
                    "h_cost": torch.empty((n_slots, S, C), dtype=self.dtype).pin_memory(),
                    "d_sample": torch.empty((n_sample_slots, C, D), dtype=self.dtype, device=dev),
                    "h_sample": torch.empty((n_sample_slots, C, D), dtype=self.dtype).pin_memory() if n_sample_slots else None,
                    "last": torch.empty(C, dtype=self.dtype, device=dev),
                    "in": [torch.cuda.Event() for _ in range(n_slots)], "done": [torch.cuda.Event() for _ in range(n_slots)],
                    "out": [torch.cuda.Event() for _ in range(n_slots)],
                    "sample_out": [torch.cuda.Event() for _ in range(n_sample_slots)], "used": [False] * n_slots,
                    "sample_used": [False] * n_sample_slots}
        s_in, s_out = pipe["s_in"], pipe["s_out"]
        epsilon = float(next(self.stepsize_schedule))
        arrays = [_native.ptr(a) for a in self._arrays()]
        sample_of_block = [-1] * len(bounds)

        def enqueue(b):
            s0, s1 = bounds[b]
            n, slot = s1 - s0, b % n_slots
            wants = bool(sample_every) and (s1 + sample_phase) % sample_every == 0     # step s1 - 1 is a sample step
            ks = -1
            with torch.cuda.device(dev):
                main = self.session.stream if self.session.stream is not None else torch.cuda.current_stream(dev)
                if pipe["used"][slot]:
                    s_in.wait_event(pipe["done"][slot])        # the block that used this slot has read its rows
                    main.wait_event(pipe["out"][slot])         # ... and its costs have left the device
                with torch.cuda.stream(s_in):
                    pipe["d_starts"][slot, :n].copy_(host_starts[s0:s1], non_blocking=True)
                    pipe["in"][slot].record(s_in)
                main.wait_event(pipe["in"][slot])
                trace = None
                if wants:
                    ks = self._host_samples % n_sample_slots
                    self._host_samples += 1
                    if pipe["sample_used"][ks]:
                        main.wait_event(pipe["sample_out"][ks])
                    trace = pipe["d_sample"][ks]
                    sample_of_block[b] = ks
                burn_left = max(0, self.burn_in_steps - self.n_iterations)
                _native.call("sgmcmc_bnn_sghmc_run_resident_f32", *arrays, _native.ptr(cf.X), _native.ptr(cf.y),
                             _native.ptr(pipe["d_starts"][slot]), None, _native.ptr(trace), None,
                             _native.ptr(pipe["d_cost"][slot]), _native.ptr(pipe["last"]), None, C, cf.n_in,
                             cf.actual_batch, float(cf.batch_size), cf.n_examples, n, min(n, burn_left),
                             int(self.burn_in_steps == 0), n if wants else 10 ** 9, epsilon, self.mdecay,
                             self.scale_grad, self._noise_seed, self.n_iterations, self.session.chain_offset,
                             main.cuda_stream)
                pipe["done"][slot].record(main)
                s_out.wait_event(pipe["done"][slot])
                with torch.cuda.stream(s_out):
                    pipe["h_cost"][slot, :n].copy_(pipe["d_cost"][slot, :n], non_blocking=True)
                    if wants:
                        pipe["h_sample"][ks].copy_(pipe["d_sample"][ks], non_blocking=True)
                        pipe["sample_out"][ks].record(s_out)
                        pipe["sample_used"][ks] = True
                    pipe["out"][slot].record(s_out)
                pipe["used"][slot] = True
            self.n_iterations += n

        queued = 0
        for b, (s0, s1) in enumerate(bounds):
            while queued < len(bounds) and queued <= b + 2:
                enqueue(queued)
                queued += 1
            slot = b % n_slots
            pipe["out"][slot].synchronize()                    # the block's costs (and sample) are in host memory
            for s in range(s0, s1):
                ks = sample_of_block[b] if s == s1 - 1 else -1
                yield (pipe["h_sample"][ks].numpy() if ks >= 0 else None), pipe["h_cost"][slot, s - s0].numpy()

    def _host_pipeline(self, depth, n_sample_slots, resident=False):
        """The native stepper of `iter_host` and its pinned result buffers, kept between calls
        (pinning [C, D] floats costs tens of milliseconds)."""
        cached = getattr(self, "_host_pipe", None)
        if cached is not None and cached[0] == (depth, n_sample_slots, resident):
            return cached[1:]
        self.close_host_pipeline()
        C, D = self.n_chains, self.n_params_per_chain
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.call("sgmcmc_bnn_host_pipeline_create", ctypes.byref(handle), C, self.cost_fun.n_in, depth,
                         int(n_sample_slots > 0) | (2 if resident else 0))
        h_cost = torch.empty((depth, C), dtype=self.dtype).pin_memory()
        h_sample = torch.empty((n_sample_slots, C, D), dtype=self.dtype).pin_memory() if n_sample_slots else None
        self._host_pipe = ((depth, n_sample_slots, resident), handle, h_cost, h_sample)
        return handle, h_cost, h_sample

    def close_host_pipeline(self):
        cached = getattr(self, "_host_pipe", None)
        if cached is not None:
            self._host_pipe = None
            _native.call("sgmcmc_bnn_host_pipeline_destroy", cached[1])

    def __del__(self):
        try:
            self.close_host_pipeline()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

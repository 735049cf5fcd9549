"""Minibatch generation -- same interface as pysgmcmc/data_batches.py:10-206, plus the
on-device index generator the B200 engine uses for many chains.

* `generate_batches` / `generate_shuffled_batches`: host generators, identical
  behaviour to the reference (NumPy ``RandomState`` stream, contiguous slices,
  ``feed_dict`` of placeholders).
* `DeviceBatchGenerator`: one MT19937 stream PER CHAIN on the GPU (kernel K7,
  csrc/mt19937.cu), producing the same ``start`` sequence
  ``RandomState(seed_j).randint(0, N - B + 1)`` bit for bit; it yields start indices
  only -- the BNN kernels slice the device-resident dataset themselves.
"""
import logging

import numpy as np
import torch

from . import _native
from .placeholders import Placeholder


def generate_batches(x, y, x_placeholder, y_placeholder, batch_size=20, seed=None):
    """Infinite generator of random minibatches (data_batches.py:10-129).

    Yields ``{x_placeholder: x[start:start+B], y_placeholder: y[start:start+B].reshape(-1, 1)}``
    with ``start = rng.randint(0, N - B + 1)``.

    >>> import numpy as np
    >>> from pysgmcmc_b200.placeholders import placeholder
    >>> N, D = 100, 3
    >>> x = np.asarray([np.random.uniform(-10, 10, D) for _ in range(N)])
    >>> y = np.asarray([np.random.choice([0., 1.]) for _ in range(N)])
    >>> xp, yp = placeholder(), placeholder()
    >>> batch_dict = next(generate_batches(x, y, xp, yp, 20))
    >>> batch_dict[xp].shape, batch_dict[yp].shape
    ((20, 3), (20, 1))
    """
    return _HostBatches(x, y, x_placeholder, y_placeholder, batch_size, seed)


def _checked_seed(seed, who):
    """The reference asserts on its arguments (tests/test_data_batches.py:78-98 expect AssertionError
    for a non-integer or non-positive batch size and for a non-integer or negative seed)."""
    assert seed is None or (isinstance(seed, int) and 0 <= seed <= 2 ** 32 - 1), \
        "%s: seed must be `None` or an integer in [0, 2**32 - 1]" % who
    return int(np.random.randint(1, 100000)) if seed is None else seed      # data_batches.py:99-100


class _HostBatches(object):
    """Iterator behind `generate_batches`: one legacy NumPy stream, one `randint` per batch -- the draw
    order `DeviceBatchGenerator` reproduces on the GPU."""

    def __init__(self, x, y, x_placeholder, y_placeholder, batch_size, seed):
        assert isinstance(batch_size, int) and batch_size > 0, \
            "generate_batches: batch size must be an integer greater than zero."
        assert y.shape[0] == x.shape[0], "Not exactly one label per datapoint!"
        self.rng = np.random.RandomState(_checked_seed(seed, "generate_batches"))
        self.x, self.y, self.keys = x, y, (x_placeholder, y_placeholder)
        self.rows = min(batch_size, x.shape[0])                    # data_batches.py:111
        if self.rows != batch_size:
            logging.error("Not enough datapoints to form a minibatch. Batchsize was set to %s", self.rows)

    def __iter__(self):
        return self

    def __next__(self):
        first = self.rng.randint(0, self.x.shape[0] - self.rows + 1)       # end-exclusive (:120)
        rows = slice(first, first + self.rows)
        return {self.keys[0]: self.x[rows], self.keys[1]: self.y[rows].reshape(-1, 1)}


def generate_shuffled_batches(x, y, x_placeholder, y_placeholder, batch_size=20, seed=None):
    """`generate_batches` with the rows of every batch shuffled, x and y by two generators in the same
    state (data_batches.py:132-206).  Like the reference this shuffles the slice VIEWS, i.e. it permutes
    the rows of the caller's x and y in place (the pairs stay matched): data_batches.py:203-205."""
    seed = _checked_seed(seed, "generate_shuffled_batches")
    shufflers = np.random.RandomState(seed), np.random.RandomState(seed)
    for batch in generate_batches(x, y, x_placeholder, y_placeholder, batch_size, seed):
        for rng, key in zip(shufflers, (x_placeholder, y_placeholder)):
            rng.shuffle(batch[key])
        yield batch


class DeviceBatchGenerator(object):
    """Per-chain minibatch start indices generated on the GPU (K7).

    Chain j draws from ``numpy.random.RandomState(seeds[j])`` exactly as
    `generate_batches` would with ``seed=seeds[j]``.  ``next(gen)`` yields
    ``{gen.starts_placeholder: int32 tensor [C]}``; `next_block(n)` returns the next
    ``n`` steps at once as ``[n, C]`` for the multi-step fused kernel.
    """

    def __init__(self, n_examples, batch_size=20, seeds=None, n_chains=None, seed=None,
                 device="cuda:0", block=256):
        assert isinstance(batch_size, int) and batch_size > 0
        if seeds is None:
            assert n_chains is not None
            if seed is None:
                seed = int(np.random.randint(1, 100000))
            # chain j continues the reference's convention "one generator per seed"
            seeds = (np.arange(n_chains, dtype=np.uint64) + np.uint64(seed)) % np.uint64(2 ** 32)
        seeds = np.asarray(seeds, dtype=np.uint64)
        assert ((0 <= seeds) & (seeds <= 2 ** 32 - 1)).all()
        self.n_examples = int(n_examples)
        self.batch_size = min(int(batch_size), self.n_examples)       # data_batches.py:111
        if self.batch_size != batch_size:
            logging.error("Not enough datapoints to form a minibatch. "
                          "Batchsize was set to %s", self.batch_size)
        self.n_chains = int(seeds.shape[0])
        self.device = torch.device(device)
        self.seeds = torch.as_tensor(seeds.astype(np.int64), device=self.device).to(torch.int32)
        self.state = torch.empty((625, self.n_chains), dtype=torch.int32, device=self.device)
        self.starts_placeholder = Placeholder("minibatch_starts")
        self.block = int(block)
        self._buf = None
        self._pos = 0
        with torch.cuda.device(self.device):
            _native.call("sgmcmc_mt19937_seed", _native.ptr(self.state), _native.ptr(self.seeds),
                         self.n_chains, _native.stream_ptr())

    def _side_stream(self):
        """A high-priority stream for K7: the index kernel is latency bound and tiny (one
        thread per chain), so its CTAs slot in next to the compute kernels of the main stream."""
        if getattr(self, "_side", None) is None:
            with torch.cuda.device(self.device):
                self._side = torch.cuda.Stream(device=self.device, priority=-1)
                # the streams were seeded on the current stream
                self._side.wait_stream(torch.cuda.current_stream(self.device))
        return self._side

    def _take_pending(self, n_steps):
        """Rows of a block that ``next(gen)`` started and did not finish: they are the next
        steps of every stream, so block requests hand them out first."""
        if self._buf is None or self._pos == self._buf.shape[0]:
            self._buf = None
            return None
        rows = self._buf[self._pos:self._pos + n_steps]
        self._pos += rows.shape[0]
        return rows

    def next_block_async(self, n_steps):
        """Start the generation of the next `n_steps` steps on the generator's side stream.
        Returns ``(starts, event)``: the consumer's stream must wait for `event` before it
        reads `starts` (int32 ``[n_steps, C]``) and call ``starts.record_stream(consumer)``.
        Blocks are generated in call order.  Rows left over from ``next(gen)`` come first
        (that case is served on the current stream)."""
        if self._buf is not None and self._pos < self._buf.shape[0]:
            out = self.next_block(n_steps)
            with torch.cuda.device(self.device):
                event = torch.cuda.Event()
                event.record(torch.cuda.current_stream(self.device))
            return out, event
        side = self._side_stream()
        with torch.cuda.device(self.device):
            if not getattr(self, "_side_owns_state", False):
                # everything queued so far on the current stream (seeding, load_state_dict,
                # earlier next_block calls) happens before the side stream touches the state
                side.wait_stream(torch.cuda.current_stream(self.device))
                self._side_owns_state = True
            # the block belongs to the SIDE stream's allocator pool: memory freed on the main
            # stream may still be read by kernels queued there, and K7 runs ahead of them
            with torch.cuda.stream(side):
                out = torch.empty((n_steps, self.n_chains), dtype=torch.int32, device=self.device)
            _native.call("sgmcmc_mt19937_starts", _native.ptr(self.state), _native.ptr(out),
                         self.n_chains, n_steps, self.n_examples - self.batch_size,
                         _native.stream_ptr(side))
            event = torch.cuda.Event()
            event.record(side)
        return out, event

    def _reclaim_state(self):
        """Make the current stream the owner of the MT19937 state again."""
        if getattr(self, "_side_owns_state", False):
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            self._side_owns_state = False

    def next_block(self, n_steps):
        """Start indices of the next `n_steps` steps: int32 ``[n_steps, C]`` (rows left over
        from a ``next(gen)`` block first, then newly generated ones)."""
        with torch.cuda.device(self.device):
            self._reclaim_state()
            head = self._take_pending(n_steps)
            n_new = n_steps - (0 if head is None else head.shape[0])
            out = torch.empty((n_new, self.n_chains), dtype=torch.int32, device=self.device)
            if n_new > 0:
                _native.call("sgmcmc_mt19937_starts", _native.ptr(self.state), _native.ptr(out),
                             self.n_chains, n_new, self.n_examples - self.batch_size,
                             _native.stream_ptr())
            if head is not None:
                out = torch.cat([head, out], dim=0) if n_new > 0 else head.contiguous()
        return out

    def state_dict(self):
        """MT19937 streams plus the not yet consumed part of the current block."""
        with torch.cuda.device(self.device):
            self._reclaim_state()
        rest = None if self._buf is None else self._buf[self._pos:].clone()
        return {"state": self.state.clone(), "pending": rest}

    def load_state_dict(self, state):
        with torch.cuda.device(self.device):
            self._reclaim_state()
        self.state.copy_(state["state"])
        self._buf = state["pending"]
        self._pos = 0
        if self._buf is not None and self._buf.shape[0] == 0:
            self._buf = None

    def __iter__(self):
        return self

    def __next__(self):
        if self._buf is None or self._pos == self._buf.shape[0]:
            self._buf = None
            self._buf = self.next_block(self.block)
            self._pos = 0
        row = self._buf[self._pos]
        self._pos += 1
        return {self.starts_placeholder: row}

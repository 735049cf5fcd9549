"""Oracle restatement of the minibatch index stream of
``pysgmcmc/data_batches.py:99-129`` (test infrastructure only).

The reference draws ``start = RandomState(seed).randint(0, N - B + 1)`` once per
step.  ``numpy.random.RandomState`` is third-party; its algorithm is restated
here (MT19937 ``init_genrand`` seeding + the legacy masked-rejection bounded
integer) and PINNED bit-exactly against NumPy itself in tests/test_oracle.py.
"""
import numpy as np

N_STATE, M_STATE = 624, 397
MATRIX_A, UPPER, LOWER = 0x9908B0DF, 0x80000000, 0x7FFFFFFF


class MT19937(object):
    def __init__(self, seed):
        # numpy legacy seeding of an int: init_genrand(seed & 0xffffffff)
        mt = [0] * N_STATE
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, N_STATE):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt = mt
        self.pos = N_STATE

    def _twist(self):
        mt = self.mt
        for i in range(N_STATE):
            y = (mt[i] & UPPER) | (mt[(i + 1) % N_STATE] & LOWER)
            mt[i] = mt[(i + M_STATE) % N_STATE] ^ (y >> 1) ^ (MATRIX_A if (y & 1) else 0)
        self.pos = 0

    def next_uint32(self):
        if self.pos >= N_STATE:
            self._twist()
        y = self.mt[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def bounded(self, max_inclusive):
        """numpy legacy ``randint(0, max_inclusive + 1)`` for max <= 2**32-1:
        AND with the smallest all-ones mask >= max, reject while > max;
        max == 0 returns 0 WITHOUT consuming a draw."""
        if max_inclusive == 0:
            return 0
        mask = max_inclusive
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        while True:
            v = self.next_uint32() & mask
            if v <= max_inclusive:
                return v


def minibatch_starts(seed, n_examples, batch_size, n_steps):
    """First `n_steps` values of ``start`` in generate_batches (data_batches.py:104-120)."""
    batch_size = min(batch_size, n_examples)                    # data_batches.py:111
    rng = MT19937(seed)
    return np.array([rng.bounded(n_examples - batch_size) for _ in range(n_steps)],
                    dtype=np.int64)

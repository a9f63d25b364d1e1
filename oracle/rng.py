"""SplitMix64 (oracle side; test infrastructure) -- identical stream in C, CUDA host and
Python so seeded inputs agree everywhere.  Seed 8349 echoes the reference's tests
(``tests/range.rs:22``)."""
import numpy as np
from .fields import R_MOD

MASK64 = (1 << 64) - 1
DEFAULT_SEED = 8349


class SplitMix64:
    def __init__(self, seed=DEFAULT_SEED):
        self.s = seed & MASK64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def fr(self):
        """uniform in [0, r): rejection-sample a 255-bit draw."""
        while True:
            v = self.next() | (self.next() << 64) | (self.next() << 128) | ((self.next() >> 1) << 192)
            if v < R_MOD:
                return v


def splitmix64_array(seed, n):
    """Vectorised: n consecutive outputs of SplitMix64(seed)."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_fr_raw_limbs(seed, n):
    """(n,4) uint64 'canonical' limbs, each value < 2^254 < r (top two bits cleared).

    Used for full-size synthetic inputs where Python-loop rejection sampling would be
    too slow; any 4-limb value below r is a valid Montgomery representative too."""
    a = splitmix64_array(seed, 4 * n).reshape(n, 4).copy()
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a

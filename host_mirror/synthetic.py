"""Seeded synthetic inputs for benchmarks: SplitMix64 so that C, CUDA-host and Python
generators agree, seed 8349 echoing the reference's tests (``tests/range.rs:22``).  Workload
generator (with ``composer.synthetic_circuit``); the test oracle keeps its own copy (``oracle/rng.py``)
so that neither side imports the other."""
import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed=8349):
        self.s = seed & _M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        v = self.s
        v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & _M64
        return v ^ (v >> 31)

    def fr(self):
        """Uniform in [0, r): rejection-sampled 255-bit draws."""
        while True:
            v = self.next() | (self.next() << 64) | (self.next() << 128) | ((self.next() >> 1) << 192)
            if v < R_MOD:
                return v


def random_fr_raw_limbs(seed, n):
    """(n, 4) uint64 limbs of values below 2^254 < r (top two bits cleared), vectorised: full-size
    scalar / NTT inputs.  Any 4-limb value below r is a valid Montgomery representative."""
    idx = np.arange(1, 4 * n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        v = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        v = (v ^ (v >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        v = (v ^ (v >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        v = v ^ (v >> np.uint64(31))
    a = v.reshape(n, 4).copy()
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a

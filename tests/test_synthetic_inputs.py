"""The product's seeded input generator (dusk-plonk_b200/synthetic.py, used by bench.py) and the
oracle's copy (oracle/rng.py) produce the same streams, so GPU and CPU arms see the same workload."""
import numpy as np

from host_mirror.synthetic import SplitMix64, random_fr_raw_limbs
from oracle import rng as orng


def test_same_streams():
    for seed in (0, 1, 8349, (1 << 64) - 1):
        a, b = SplitMix64(seed), orng.SplitMix64(seed)
        assert [a.next() for _ in range(100)] == [b.next() for _ in range(100)]
        a, b = SplitMix64(seed), orng.SplitMix64(seed)
        assert [a.fr() for _ in range(40)] == [b.fr() for _ in range(40)]
    for seed, n in ((8349, 1), (8350, 1000), (4242, 4097)):
        assert np.array_equal(random_fr_raw_limbs(seed, n), orng.random_fr_raw_limbs(seed, n))

"""Times zkp_srs_generate (fixed-base powers + the window table every commit uses; PlonkParams::setup / trim)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np
import dusk_plonk_b200 as z
from host_mirror.synthetic import random_fr_raw_limbs
ctx = z.Context(0)
tau = random_fr_raw_limbs(4242, 1)[0]
for k in (16, 20):
    n = (1 << k) + 7
    t0 = time.perf_counter(); srs = ctx.srs_generate(tau, n); ctx.sync(); dt = time.perf_counter() - t0
    print("srs_generate + window table, 2^%d points: %.1f ms" % (k, dt * 1e3))
    srs.free()

"""The device arithmetic source (csrc/arith.cuh, csrc/g1.cuh) compiled for the host with
the PTX carry flag emulated: the exact even/odd-accumulator Montgomery algorithm and the
XYZZ formulas the kernels run, checked against the big-integer oracle without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import curve
from oracle.fields import (FQ_MONT_R, FR_MONT_R, P_MOD, R_MOD, fq_from_mont_limbs, fq_to_mont_limbs,
                           fr_from_mont_limbs, fr_from_raw_limbs, fr_to_mont_limbs, fr_to_raw_limbs,
                           g1_from_mont_limbs, g1_to_mont_limbs)
from oracle.rng import SplitMix64

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dusk-plonk_b200")


@pytest.fixture(scope="module")
def L():
    so = os.path.join(PKG, "_build", "libhost_arith_test.so")
    subprocess.check_call(["make", "-C", PKG, "-s", "_build/libhost_arith_test.so"])
    return ctypes.CDLL(so)


def p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def call2(L, fn, a, b, n):
    o = np.zeros(n, dtype=np.uint64)
    getattr(L, fn)(p(a), p(b), p(o))
    return o


def call1(L, fn, a, n):
    o = np.zeros(n, dtype=np.uint64)
    getattr(L, fn)(p(a), p(o))
    return o


def test_fr(L):
    rng = SplitMix64(1)
    vals = [0, 1, 2, R_MOD - 1, R_MOD - 2, (1 << 255) % R_MOD, FR_MONT_R, R_MOD - FR_MONT_R]
    vals += [rng.fr() for _ in range(200)]
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        al, bl = fr_to_mont_limbs([a])[0], fr_to_mont_limbs([b])[0]
        assert fr_from_mont_limbs(call2(L, "ht_fr_mul", al, bl, 4))[0] == a * b % R_MOD
        assert fr_from_mont_limbs(call2(L, "ht_fr_add", al, bl, 4))[0] == (a + b) % R_MOD
        assert fr_from_mont_limbs(call2(L, "ht_fr_sub", al, bl, 4))[0] == (a - b) % R_MOD
        assert fr_from_mont_limbs(call1(L, "ht_fr_neg", al, 4))[0] == (-a) % R_MOD
        assert fr_from_raw_limbs(call1(L, "ht_fr_from_mont", al, 4))[0] == a
        assert fr_from_mont_limbs(call1(L, "ht_fr_to_mont", fr_to_raw_limbs([a])[0], 4))[0] == a
    for a in vals[:12]:
        al = fr_to_mont_limbs([a])[0]
        assert fr_from_mont_limbs(call1(L, "ht_fr_inv", al, 4))[0] == (pow(a, -1, R_MOD) if a else 0)


def test_fq(L):
    rng = SplitMix64(2)
    vals = [0, 1, 2, P_MOD - 1, P_MOD - 2, FQ_MONT_R] + [(rng.fr() * rng.fr() * rng.next()) % P_MOD for _ in range(200)]
    for i, a in enumerate(vals):
        b = vals[(i * 5 + 1) % len(vals)]
        al, bl = fq_to_mont_limbs([a])[0], fq_to_mont_limbs([b])[0]
        assert fq_from_mont_limbs(call2(L, "ht_fq_mul", al, bl, 6))[0] == a * b % P_MOD
        assert fq_from_mont_limbs(call2(L, "ht_fq_add", al, bl, 6))[0] == (a + b) % P_MOD
        assert fq_from_mont_limbs(call2(L, "ht_fq_sub", al, bl, 6))[0] == (a - b) % P_MOD
        assert fq_from_mont_limbs(call1(L, "ht_fq_neg", al, 6))[0] == (-a) % P_MOD
    for a in vals[:8]:
        assert fq_from_mont_limbs(call1(L, "ht_fq_inv", fq_to_mont_limbs([a])[0], 6))[0] == (pow(a, -1, P_MOD) if a else 0)


def test_xyzz_group_law_with_edge_cases(L):
    G = curve.G1_GEN

    def aff(pt):
        return g1_to_mont_limbs([pt])[0].copy()

    def toaff(acc):
        o = np.zeros(12, dtype=np.uint64)
        L.ht_xyzz_to_affine(p(acc), p(o))
        return g1_from_mont_limbs(o)[0]

    acc = np.zeros(24, dtype=np.uint64)
    assert toaff(acc) is None
    exp = None
    for i, pt in enumerate([curve.mul(G, k) for k in (5, 5, 5, 11, 3)]):  # repeated point -> doubling branch
        L.ht_xyzz_madd(p(acc), p(aff(pt)), 0)
        exp = curve.add(exp, pt)
        assert toaff(acc) == exp, i
    L.ht_xyzz_madd(p(acc), p(aff(exp)), 1)  # P + (-P) -> infinity
    assert toaff(acc) is None
    L.ht_xyzz_madd(p(acc), p(aff(None)), 0)  # affine infinity is skipped
    assert toaff(acc) is None
    L.ht_xyzz_madd(p(acc), p(aff(G)), 1)
    assert toaff(acc) == curve.neg(G)
    a2 = np.zeros(24, dtype=np.uint64)
    L.ht_xyzz_madd(p(a2), p(aff(curve.mul(G, 7))), 0)
    L.ht_xyzz_madd(p(a2), p(aff(curve.mul(G, 9))), 0)
    L.ht_xyzz_add(p(acc), p(a2))
    assert toaff(acc) == curve.mul(G, 15)
    L.ht_xyzz_add(p(acc), p(acc.copy()))  # equal XYZZ points -> doubling branch
    assert toaff(acc) == curve.mul(G, 30)
    L.ht_xyzz_dbl(p(acc))
    assert toaff(acc) == curve.mul(G, 60)
    a3 = np.zeros(24, dtype=np.uint64)
    L.ht_xyzz_madd(p(a3), p(aff(curve.mul(G, 60))), 1)
    L.ht_xyzz_add(p(acc), p(a3))
    assert toaff(acc) is None
    L.ht_xyzz_add(p(acc), p(a3))
    assert toaff(acc) == curve.mul(G, R_MOD - 60)


def test_binary_gcd_inversion_matches_big_integer_inverse(L):
    """csrc/host_inv.h (the host inversion at every transcript synchronisation point): y^-1 mod m for both
    moduli, edge values (1, m - 1, powers of two around the 31 / 64-bit approximation boundaries, short
    values) and random ones; 0 maps to 0 like the Fermat chain it replaces."""
    import random
    rng = random.Random(3)

    def limbs(v, n):
        return np.array([(v >> (64 * i)) & ((1 << 64) - 1) for i in range(n)], dtype=np.uint64)

    def val(a):
        return sum(int(x) << (64 * i) for i, x in enumerate(a))

    for m, n in ((R_MOD, 4), (P_MOD, 6)):
        mm = limbs(m, n)
        cases = [1, 2, 3, m - 1, m - 2, (m - 1) // 2, (m + 1) // 2, 2 ** 31, 2 ** 31 - 1, 2 ** 33, 2 ** 62, 2 ** 64 - 1,
                 2 ** 64, 2 ** 65 + 1, 2 ** 127, 2 ** 128 + 5]
        cases += [rng.randrange(1, m) for _ in range(3000)]
        cases += [(rng.randrange(1, 2 ** rng.randrange(1, 64 * n)) % m) or 1 for _ in range(1000)]
        for y in cases:
            out = np.zeros(n, dtype=np.uint64)
            L.ht_inv_mod(p(limbs(y, n)), p(mm), p(out), n)
            assert val(out) == pow(y, -1, m), hex(y)
        out = np.ones(n, dtype=np.uint64)
        L.ht_inv_mod(p(limbs(0, n)), p(mm), p(out), n)
        assert val(out) == 0

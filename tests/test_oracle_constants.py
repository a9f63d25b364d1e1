"""Pins the oracle to every known-answer the reference holds in-tree for this path
(SURVEY 8c): MINUS_ONE limbs, K1..K3, domain ordering, plus public BLS12-381 facts."""
import numpy as np

from oracle import curve, ntt
from oracle.fields import (FQ_INV32, FQ_INV64, FR_INV32, FR_INV64, K1, K2, K3, P_MOD, R_MOD,
                           ROOT_OF_UNITY, MULTIPLICATIVE_GENERATOR, fr_to_mont_limbs,
                           fr_from_mont_limbs, domain_generator)


def test_minus_one_limbs_src_lib_rs_583():
    # /root/reference/src/lib.rs:583-588
    expect = [0xfffffffd00000003, 0xfb38ec08fffb13fc, 0x99ad88181ce5880f, 0x5bc8f5f97cd877d8]
    got = [int(x) for x in fr_to_mont_limbs([R_MOD - 1])[0]]
    assert got == expect
    assert fr_from_mont_limbs(np.array([expect], dtype=np.uint64)) == [R_MOD - 1]


def test_permutation_coset_constants():
    # /root/reference/src/permutation.rs:28-30
    assert (K1, K2, K3) == (7, 13, 17)
    # k_i H must be distinct cosets of the 2^32 subgroup: (k_i/k_j)^(2^32) != 1
    ks = [1, K1, K2, K3]
    for i in range(4):
        for j in range(i):
            q = ks[i] * pow(ks[j], -1, R_MOD) % R_MOD
            assert pow(q, 1 << 32, R_MOD) != 1


def test_field_parameters():
    assert R_MOD.bit_length() == 255 and P_MOD.bit_length() == 381
    assert (R_MOD - 1) % (1 << 32) == 0 and (R_MOD - 1) % (1 << 33) != 0
    assert ROOT_OF_UNITY == 0x16a2a19edfe81f20d09b681922c813b4b63683508c2280b93829971f439f0d2b
    assert pow(ROOT_OF_UNITY, 1 << 32, R_MOD) == 1 and pow(ROOT_OF_UNITY, 1 << 31, R_MOD) == R_MOD - 1
    # 7 is a quadratic non-residue, hence a valid coset shift / multiplicative generator
    assert pow(MULTIPLICATIVE_GENERATOR, (R_MOD - 1) // 2, R_MOD) == R_MOD - 1
    assert (FR_INV32, FQ_INV32) == (0xffffffff, 0xfffcfffd)
    assert (FR_INV64, FQ_INV64) == (0xfffffffeffffffff, 0x89f3fffcfffcfffd)


def test_domain_ordering_elements_i_is_w_pow_i():
    # src/permutation.rs:148-166,764-772: roots[index] == w^index; :1036-1038: last * w == 1
    f = ntt.Fft(4)
    w = f.generator()
    assert f.elements[2] == pow(w, 2, R_MOD) and f.elements[3] == pow(w, 3, R_MOD)
    assert f.elements[-1] * w % R_MOD == 1
    assert domain_generator(0) == 1 and domain_generator(1) == R_MOD - 1


def test_g1_generator_and_order():
    assert curve.is_on_curve(curve.G1_GEN)
    assert curve.mul(curve.G1_GEN, R_MOD - 1) == curve.neg(curve.G1_GEN)
    assert curve.add(curve.mul(curve.G1_GEN, R_MOD - 1), curve.G1_GEN) is None


def test_dft_matches_definition_and_roundtrip():
    from oracle.rng import SplitMix64
    rng = SplitMix64()
    for k in (0, 1, 3, 5):
        f = ntt.Fft(k)
        v = [rng.fr() for _ in range(max(1, (1 << k) - 2))]
        assert f.dft(v) == ntt.dft_naive(v, k)
        assert f.idft(f.dft(v))[:len(v)] == v
        c = f.coset_dft(v)
        assert c == [ntt.poly_eval(v, 7 * pow(f.w, j, R_MOD) % R_MOD) for j in range(1 << k)]
        assert f.coset_idft(c)[:len(v)] == v
    # vanishing polynomial on the 8n coset takes 8 distinct values
    f8 = ntt.Fft(6)
    zh = f8.compute_vanishing_poly_over_coset(8)
    assert zh[:8] == zh[8:16] and len(set(zh)) == 8
    assert zh[3] == (pow(7 * pow(f8.w, 3, R_MOD), 8, R_MOD) - 1) % R_MOD


def test_msm_algorithms_agree():
    from oracle.rng import SplitMix64
    rng = SplitMix64(7)
    tau = rng.fr()
    n = 40
    pts = curve.srs_powers(tau, n)
    sc = [rng.fr() for _ in range(n)]
    sc[0], sc[1], sc[2] = 0, 1, R_MOD - 1
    a = curve.msm_naive(pts, sc)
    assert a == curve.msm_pippenger(pts, sc) == curve.msm_pippenger(pts, sc, c=5)
    assert a == curve.commit_known_dlog([pow(tau, i, R_MOD) for i in range(n)], sc)

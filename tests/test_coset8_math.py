"""CPU: the decomposition the multi-GPU prover rests on (csrc/ntt.cu "cosets of the n-domain", csrc/create_proof.cu
coset8_combine / openings_sharded), restated with the oracle's big-integer transforms and checked against the
reference's 8n-point coset transforms (src/prover/quotient_poly.rs:54-58,115) for G = 1, 2, 4, 8 simulated ranks:
coset-wise forward values, the one-exchange inverse with its 8 x 8 combination and slab ownership, the slab-wise
openings and the carried division by (X - z)."""
import pytest

from oracle.fields import R_MOD, domain_generator
from oracle.ntt import Fft, poly_eval
from oracle.rng import SplitMix64

r = R_MOD
g = 7


def coset_values(coeffs, k, u):
    """values of the polynomial on coset u: n-point transform of the folded, h_u^e-scaled coefficients"""
    n = 1 << k
    w8 = domain_generator(k + 3)
    h = g * pow(w8, u, r) % r
    fold = pow(h, n, r)
    c = [0] * n
    for e, v in enumerate(coeffs):
        c[e % n] = (c[e % n] + v * pow(fold, e // n, r)) % r
    return Fft(k).dft([c[e] * pow(h, e, r) % r for e in range(n)])


@pytest.mark.parametrize("k", [3, 5])
def test_forward_cosets_interleave_to_the_8n_coset_dft(k):
    n = 1 << k
    rng = SplitMix64(k)
    poly = [rng.fr() for _ in range(n + 3)]
    want = Fft(k + 3).coset_dft(poly)
    for u in range(8):
        assert coset_values(poly, k, u) == [want[8 * m + u] for m in range(n)]


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_one_exchange_inverse_and_slab_openings(G):
    k = 6          # slab = n / G >= 8: the product takes the slab-wise path only then (create_proof.cu)
    n = 1 << k
    rng = SplitMix64(100 + G)
    t = [rng.fr() for _ in range(4 * n + 7)] + [0] * (4 * n - 7)          # deg t = 4n + 6
    T = Fft(k + 3).coset_dft(t)
    w8 = domain_generator(k + 3)
    fn = Fft(k)
    nloc, slab = 8 // G, n // G
    # every rank: n-point inverses of its cosets, scaled by h_u^-e / 8
    inv8 = pow(8, -1, r)
    Y = {}
    for rank in range(G):
        for u in range(rank * nloc, (rank + 1) * nloc):
            h_inv = pow(g * pow(w8, u, r) % r, -1, r)
            B = fn.idft([T[8 * m + u] for m in range(n)])
            Y[u] = [B[e] * pow(h_inv, e, r) % r * inv8 % r for e in range(n)]
    # the exchange: rank s receives slab s of all eight; combination c[q][u] = g^(-n q) w_8^(-u q)
    w8_8 = domain_generator(3)
    gni = pow(pow(g, n, r), -1, r)
    slabs = {}
    for s in range(G):
        for q in range(8):
            slabs[(s, q)] = [sum(pow(gni, q, r) * pow(w8_8, (-u * q) % 8, r) % r * Y[u][s * slab + i] for u in range(8)) % r
                             for i in range(slab)]
    for s in range(G):
        for q in range(8):
            assert slabs[(s, q)] == t[q * n + s * slab: q * n + (s + 1) * slab], (s, q)
    # slab-wise opening of t at z: sum_s z^lo_s sum_q z^(q n) slab_q(z); the tail (index >= n of t_4) sits with the last rank
    z = rng.fr()
    total = 0
    for s in range(G):
        for q in range(4):
            piece = list(slabs[(s, q)])
            if q == 3 and s == G - 1:
                piece += slabs[(0, 4)][:7]
            total = (total + pow(z, q * n + s * slab, r) * poly_eval(piece, z)) % r
    assert total == poly_eval(t, z)
    # carried division of a polynomial of n + 7 coefficients by (X - z), rank ranges [s slab, (s + 1) slab) + tail
    a = [rng.fr() for _ in range(n + 7)]
    wfull = [0] * (n + 6)
    acc = 0
    for kk in range(n + 6, 0, -1):
        acc = (a[kk] + z * acc) % r
        wfull[kk - 1] = acc
    ranges = [(s * slab, (s + 1) * slab if s < G - 1 else n + 7) for s in range(G)]
    P = [poly_eval(a[lo:hi], z) for lo, hi in ranges]
    for s, (lo, hi) in enumerate(ranges):
        carry = 0
        for s2 in range(G - 1, s, -1):
            carry = (P[s2] + pow(z, ranges[s2][1] - ranges[s2][0], r) * carry) % r
        arr = [0] + a[lo + 1:hi] + ([carry] if s < G - 1 else [])
        out, acc = [0] * (len(arr) - 1), 0
        for j in range(len(arr) - 1, 0, -1):
            acc = (arr[j] + z * acc) % r
            out[j - 1] = acc
        assert out == wfull[lo:lo + len(out)], s

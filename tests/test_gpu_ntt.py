"""GPU parity: the CUDA NTT family (through the C ABI) against the oracle.

Mirrors how the reference calls ``poly_commit::Fft`` (src/prover.rs:121-124,192,229;
src/prover/quotient_poly.rs:54-58,115; src/key.rs:121-131,226-245): by-value transforms,
zero-padded short inputs, natural-order output.  Bit-exact (integer arithmetic)."""
import numpy as np
import pytest

from oracle import ntt as ontt
from oracle.fields import R_MOD, fr_from_mont_limbs, fr_to_mont_limbs
from oracle.rng import SplitMix64, random_fr_raw_limbs

pytestmark = pytest.mark.gpu

MODES = [(False, False), (True, False), (False, True), (True, True)]


def _oracle(f, v, inverse, coset):
    return {(False, False): f.dft, (True, False): f.idft, (False, True): f.coset_dft,
            (True, True): f.coset_idft}[(inverse, coset)](v)


@pytest.mark.parametrize("k", list(range(0, 13)))
def test_small_sizes_exhaustive_vs_python(ctx, k):
    """2^0 .. 2^12 (config 1 uses n = 2^9, 8n = 2^12), every mode, ragged input lengths."""
    rng = SplitMix64(8349 + k)
    n = 1 << k
    f = ontt.Fft(k)
    for m in sorted({n, max(1, n - 3), max(1, n // 8 + 3 if n >= 8 else 1), 1}):
        if m > n:
            continue
        v = [rng.fr() for _ in range(m)]
        vm = fr_to_mont_limbs(v)
        for inverse, coset in MODES:
            got = fr_from_mont_limbs(ctx.ntt(vm, k, inverse, coset))
            assert got == _oracle(f, v, inverse, coset), (k, m, inverse, coset)


@pytest.mark.parametrize("k", [13, 16, 17, 19, 20])
def test_mid_sizes_vs_c_oracle(ctx, cport, k):
    """Pass plans with 2 and 3 digits, against the C restatement on full vectors."""
    n = 1 << k
    data = random_fr_raw_limbs(8349 + k, n)
    for inverse, coset in MODES:
        got = ctx.ntt(data, k, inverse, coset)
        exp = cport.ntt(data, k, inverse, coset)
        assert np.array_equal(got, exp), (k, inverse, coset)
    # zero-padded short input (n + 3 blinded coefficients into 8n, quotient_poly.rs:54-58)
    m = n // 8 + 3
    got = ctx.ntt(data[:m], k, False, True)
    exp = cport.ntt(data[:m], k, False, True)
    assert np.array_equal(got, exp)


def test_edge_vectors(ctx):
    k = 10
    n = 1 << k
    zero = np.zeros((n, 4), dtype=np.uint64)
    for inverse, coset in MODES:
        assert not ctx.ntt(zero, k, inverse, coset).any()
    one = fr_to_mont_limbs([1])
    # dft of the constant polynomial 1 is all ones; of X it is elements[i] = w^i
    ev = fr_from_mont_limbs(ctx.ntt(one, k))
    assert ev == [1] * n
    x = fr_to_mont_limbs([0, 1])
    f = ontt.Fft(k)
    assert fr_from_mont_limbs(ctx.ntt(x, k)) == f.elements
    # coset_dft(X) = linear_evaluations of src/key.rs:223-245: g * w^i
    assert fr_from_mont_limbs(ctx.ntt(x, k, False, True)) == [7 * e % R_MOD for e in f.elements]
    # maximal elements r - 1
    top = fr_to_mont_limbs([R_MOD - 1] * n)
    assert fr_from_mont_limbs(ctx.ntt(top, k)) == f.dft([R_MOD - 1] * n)


def test_fft_elements_and_constants(ctx):
    import dusk_plonk_b200 as z
    f = z.Fft(ctx, 9)
    o = ontt.Fft(9)
    assert fr_from_mont_limbs(f.elements) == o.elements
    assert fr_from_mont_limbs(f.generator()) == [o.w]
    assert fr_from_mont_limbs(f.size_inv()) == [o.n_inv]


def test_device_resident_and_batched(ctx, cport):
    k = 14
    n = 1 << k
    batch = 4  # the four wire polynomials of src/prover.rs:121-124
    data = random_fr_raw_limbs(99, n * batch)
    src = ctx.upload(data)
    dst = ctx.alloc(n * batch)
    ctx.ntt_dev_batch(src, n, n, dst, n, k, True, False, batch)
    got = dst.download()
    for b in range(batch):
        assert np.array_equal(got[b * n:(b + 1) * n], cport.ntt(data[b * n:(b + 1) * n], k, inverse=True)), b
    # in place, single polynomial
    one = ctx.upload(data[:n])
    ctx.ntt_dev(one, n, one, k, False, True)
    assert np.array_equal(one.download(), cport.ntt(data[:n], k, coset=True))


@pytest.mark.parametrize("k", [22, 24, 26])
def test_full_size_properties(ctx, k):
    """BASELINE sizes: round trip, spot evaluation against Horner, and linearity -- the
    size-independent properties of an exact DFT."""
    n = 1 << k
    a = random_fr_raw_limbs(1000 + k, n)
    da = ctx.upload(a)
    ev = ctx.alloc(n)
    ctx.ntt_dev(da, n, ev, k, False, False)
    evh = ev.download()
    # spot-check out[j] = sum_i a_i w^(ij) with a sparse polynomial so Horner is cheap:
    sp = np.zeros((n, 4), dtype=np.uint64)
    idxs = [0, 1, 5, n // 3, n // 2 + 1, n - 1]
    rng = SplitMix64(k)
    vals = [rng.fr() for _ in idxs]
    sp[idxs] = fr_to_mont_limbs(vals)
    dsp = ctx.upload(sp)
    evs = ctx.alloc(n)
    ctx.ntt_dev(dsp, n, evs, k, False, False)
    evsh = evs.download()
    w = ontt.Fft(k).w
    for j in [0, 1, 2, n // 2, n - 1, 123457 % n]:
        x = pow(w, j, R_MOD)
        exp = sum(v * pow(x, i, R_MOD) for i, v in zip(idxs, vals)) % R_MOD
        assert fr_from_mont_limbs(evsh[j:j + 1]) == [exp], j
    # linearity: dft(a + sp) == dft(a) + dft(sp) at sampled positions
    summ = fr_to_mont_limbs([(x + y) % R_MOD for x, y in zip(fr_from_mont_limbs(a[idxs]), vals)])
    a2 = a.copy()
    a2[idxs] = summ
    d2 = ctx.upload(a2)
    ctx.ntt_dev(d2, n, d2, k, False, False)
    e2 = d2.download()
    pos = [0, 7, n // 2, n - 1]
    lhs = fr_from_mont_limbs(e2[pos])
    rhs = [(x + y) % R_MOD for x, y in zip(fr_from_mont_limbs(evh[pos]), fr_from_mont_limbs(evsh[pos]))]
    assert lhs == rhs
    # round trips, plain and coset
    ctx.ntt_dev(ev, n, ev, k, True, False)
    assert np.array_equal(ev.download(), a)
    ctx.ntt_dev(da, n, ev, k, False, True)
    ctx.ntt_dev(ev, n, ev, k, True, True)
    assert np.array_equal(ev.download(), a)

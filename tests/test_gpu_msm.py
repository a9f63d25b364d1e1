"""GPU parity: the CUDA Pippenger MSM / KZG commit (through the C ABI) against the oracle.

Mirrors ``keypair.commit(&poly)?`` (src/prover.rs:133-136,194,262-265,440,452) and
``.unwrap_or_default()`` (src/key.rs:138-154): affine result, Err on degree overflow,
identity for the zero polynomial, trailing zeros ignored.  Bit-exact."""
import numpy as np
import pytest

from oracle import curve
from oracle.fields import (R_MOD, fr_from_mont_limbs, fr_to_mont_limbs, fr_to_raw_limbs, g1_from_mont_limbs,
                           g1_to_mont_limbs)
from oracle.rng import SplitMix64, random_fr_raw_limbs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_srs():
    rng = SplitMix64(8349)
    tau = rng.fr()
    n = 300
    return tau, curve.srs_powers(tau, n)


def test_srs_generation_matches_oracle(ctx, small_srs):
    tau, pts = small_srs
    srs = ctx.srs_generate(fr_to_mont_limbs([tau])[0], len(pts))
    assert g1_from_mont_limbs(srs.download()) == pts


@pytest.mark.parametrize("n", [1, 2, 3, 31, 64, 257, 300])
def test_small_msm_vs_python(ctx, small_srs, n):
    tau, pts = small_srs
    srs = ctx.srs_load(g1_to_mont_limbs(pts))
    rng = SplitMix64(n)
    sc = [rng.fr() for _ in range(n)]
    if n > 3:
        sc[0], sc[1], sc[2], sc[3] = 0, 1, R_MOD - 1, (1 << 254) % R_MOD
    exp = curve.commit_known_dlog([pow(tau, i, R_MOD) for i in range(n)], sc)
    assert g1_from_mont_limbs(ctx.msm(srs, fr_to_mont_limbs(sc)))[0] == exp
    if n <= 64:
        assert exp == curve.msm_naive(pts[:n], sc)


@pytest.mark.parametrize("c", [2, 3, 5, 8, 11, 13, 16])
def test_every_window_size(ctx, small_srs, c):
    """Signed-digit decomposition and carry handling for many window widths."""
    tau, pts = small_srs
    rng = SplitMix64(100 + c)
    n = 200
    sc = [rng.fr() for _ in range(n)]
    # digits that hit the +2^(c-1) boundary, all-ones windows and the largest scalar
    sc[:6] = [R_MOD - 1, (1 << (c - 1)), (1 << c) - 1, (1 << 255) % R_MOD, ((1 << 254) - 1), (1 << c) + (1 << (c - 1))]
    ctx.set_msm_window(c)   # baked into the SRS window table at load time
    try:
        srs = ctx.srs_load(g1_to_mont_limbs(pts))
    finally:
        ctx.set_msm_window(0)
    got = g1_from_mont_limbs(ctx.msm(srs, fr_to_mont_limbs(sc)))[0]
    assert got == curve.commit_known_dlog([pow(tau, i, R_MOD) for i in range(n)], sc)


@pytest.mark.parametrize("profile", ["all_equal", "tiny_values", "carry_only_top", "one_hot"])
def test_skewed_digit_distributions(ctx, profile):
    """Giant buckets (every scalar sharing a digit, small selector-like coefficients, a top
    window holding only the carry) are cut into tasks and merged; result unchanged."""
    n = 3000
    rng = SplitMix64(7)
    tau = rng.fr()
    ctx.set_msm_window(15 if profile == "carry_only_top" else 0)
    try:
        srs = ctx.srs_generate(fr_to_mont_limbs([tau])[0], n)
    finally:
        ctx.set_msm_window(0)
    if profile == "all_equal":
        sc = [R_MOD - 5] * n
    elif profile == "tiny_values":
        sc = [(i * 7) % 3 for i in range(n)]
    elif profile == "carry_only_top":
        sc = [R_MOD - 1 - i for i in range(n)]
    else:
        sc = [0] * n
        sc[n - 1] = 12345
    exp = curve.commit_known_dlog([pow(tau, i, R_MOD) for i in range(n)], sc)
    assert g1_from_mont_limbs(ctx.msm(srs, fr_to_mont_limbs(sc)))[0] == exp


def test_degenerate_bases_and_cancellation(ctx):
    G = curve.G1_GEN
    g = g1_to_mont_limbs([G])[0]
    bases = np.tile(g, (64, 1))
    bases[10] = 0  # point at infinity in the SRS
    srs = ctx.srs_load(bases)
    ones = fr_to_mont_limbs([1] * 64)
    assert g1_from_mont_limbs(ctx.msm(srs, ones))[0] == curve.mul(G, 63)  # same bucket, doubling branch
    mixed = fr_to_mont_limbs([5] * 32 + [R_MOD - 5] * 32)
    bases2 = np.tile(g, (64, 1))
    srs2 = ctx.srs_load(bases2)
    assert g1_from_mont_limbs(ctx.msm(srs2, mixed))[0] is None  # total cancels to the identity
    assert g1_from_mont_limbs(ctx.msm(srs2, fr_to_mont_limbs([0] * 64)))[0] is None
    assert g1_from_mont_limbs(ctx.msm(srs2, fr_to_mont_limbs([R_MOD - 1] * 3)))[0] == curve.mul(G, R_MOD - 3)


def test_commit_semantics(ctx, small_srs):
    """PlonkParams::commit: degree check ignores trailing zeros; overflow is an error;
    the zero polynomial commits to the identity (src/key.rs:138-154)."""
    import dusk_plonk_b200 as z
    tau, pts = small_srs
    pp = z.PlonkParams.from_points(ctx, g1_to_mont_limbs(pts[:40]))
    rng = SplitMix64(5)
    coeffs = [rng.fr() for _ in range(40)]
    exp = curve.commit_known_dlog([pow(tau, i, R_MOD) for i in range(40)], coeffs)
    assert g1_from_mont_limbs(pp.commit(z.Coefficients(fr_to_mont_limbs(coeffs))).xy)[0] == exp
    # 5n-long vector with only the first 40 non-zero (t_4 = t_poly[3n..], src/prover.rs:259)
    padded = coeffs + [0] * 160
    assert g1_from_mont_limbs(pp.commit(z.Coefficients(fr_to_mont_limbs(padded))).xy)[0] == exp
    too_long = coeffs + [0] * 5 + [1]
    with pytest.raises(z.Error):
        pp.commit(z.Coefficients(fr_to_mont_limbs(too_long)))
    assert pp.commit_or_default(z.Coefficients(fr_to_mont_limbs(too_long))).is_identity()
    assert pp.commit(z.Coefficients(fr_to_mont_limbs([0] * 17))).is_identity()
    # device-resident coefficients
    buf = ctx.upload(fr_to_mont_limbs(padded))
    assert g1_from_mont_limbs(pp.commit(buf).xy)[0] == exp


@pytest.mark.parametrize("logn", [12, 16])
def test_mid_size_vs_c_oracle(ctx, cport, logn):
    n = 1 << logn
    tau = fr_to_mont_limbs([SplitMix64(logn).fr()])[0]
    srs = ctx.srs_generate(tau, n)
    bases = srs.download()
    sc = random_fr_raw_limbs(77 + logn, n)
    got = ctx.msm(srs, sc)
    exp = cport.msm_g1(bases, sc)
    assert np.array_equal(got, exp)
    # PLONK-like skewed scalars: 30 % zero, 30 % < 2^8, rest uniform
    sk = sc.copy()
    sel = np.arange(n) % 10
    sk[sel < 3] = 0
    small = fr_to_mont_limbs(list(range(1, 257)))
    idx = np.nonzero((sel >= 3) & (sel < 6))[0]
    sk[idx] = small[idx % 256]
    assert np.array_equal(ctx.msm(srs, sk), cport.msm_g1(bases, sk))


@pytest.mark.parametrize("logn", [20, 24])
def test_full_size_known_dlog(ctx, logn):
    """BASELINE sizes: bases are [tau^i]G with known tau, so the exact answer is
    (sum_i s_i tau^i mod r) * G -- O(N) field work on the host."""
    n = 1 << logn
    tau_i = SplitMix64(4242).fr()
    srs = ctx.srs_generate(fr_to_mont_limbs([tau_i])[0], n)
    sc_mont = random_fr_raw_limbs(31337 + logn, n)
    got = g1_from_mont_limbs(ctx.msm(srs, sc_mont))[0]
    # scalars as canonical integers: the limbs are Montgomery representatives
    from oracle.fields import FR_MONT_RINV, _from_limbs_fast
    acc, t = 0, 1
    for v in _from_limbs_fast(sc_mont, 4):
        acc = (acc + v * t) % R_MOD
        t = t * tau_i % R_MOD
    acc = acc * FR_MONT_RINV % R_MOD
    assert got == curve.mul(curve.G1_GEN, acc)
    # spot-check the generated bases themselves
    for i in (0, 1, n // 2, n - 1):
        assert g1_from_mont_limbs(srs.download(i, 1))[0] == curve.mul(curve.G1_GEN, pow(tau_i, i, R_MOD))


def test_commit_batch_matches_single_commits(ctx):
    """The batched launch sequence (wire / quotient-chunk commits) equals separate commits,
    including ragged lengths, a zero polynomial, trailing zeros past the SRS and one overflow."""
    from dusk_plonk_b200.plonk_params import PlonkParams, Error
    n = 700
    rng = SplitMix64(21)
    tau = rng.fr()
    pp = PlonkParams(ctx, ctx.srs_generate(fr_to_mont_limbs([tau])[0], n))
    polys = [[rng.fr() for _ in range(n)], [rng.fr() for _ in range(33)], [0] * 50,
             [rng.fr() for _ in range(n - 5)] + [0] * 400, [5], [rng.fr() % 4 for _ in range(n)]]
    bufs = [ctx.upload(fr_to_mont_limbs(p)) for p in polys]
    got = pp.commit_batch(bufs)
    for p, b, g in zip(polys, bufs, got):
        exp = curve.commit_known_dlog([pow(tau, i, R_MOD) for i in range(len(p))], p) if any(p) else None
        assert g.affine() == exp
        assert pp.commit(b) == g
    # views into one buffer, like the quotient chunks t_low .. t_4
    big = ctx.upload(fr_to_mont_limbs(polys[0] + polys[3]))
    from dusk_plonk_b200.ffi import BufferView
    v = pp.commit_batch([BufferView(big, 0, n), BufferView(big, n, len(polys[3]))])
    assert v[0] == got[0] and v[1] == got[3]
    bad = ctx.upload(fr_to_mont_limbs([1] * (n + 1)))
    with pytest.raises(Error):
        pp.commit_batch([bufs[0], bad])
    with pytest.raises(Error):
        pp.commit(bad)

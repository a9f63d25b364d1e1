"""GPU parity: device-resident prover rounds (csrc/prover.cu through the C ABI) and the full
``create_proof`` against the CPU restatement (oracle/plonk.py), bit-exact.

Mirrors the reference's integration tests (tests/range.rs, tests/logic.rs, tests/ecc.rs,
README TestCircuit): compile -> create_proof -> verify; unsatisfied circuits make
create_proof return Err.  Blinders, SRS trapdoor and transcript are shared explicit inputs."""
import os

import numpy as np
import pytest

import dusk_plonk_b200 as z
from host_mirror.composer import SynthesizedCircuit
from oracle.merlin import Transcript as OTranscript
from dusk_plonk_b200.field import fr_from_mont, fr_to_mont, fr_to_mont1
from dusk_plonk_b200.plonk_params import Error, PlonkParams
from oracle import plonk as oplonk
from oracle.fields import K1, K2, K3, R_MOD, domain_generator
from oracle.ntt import Fft as OFft, poly_eval
from oracle.rng import SplitMix64

import circuits

pytestmark = pytest.mark.gpu


def rand_vec(seed, n):
    rng = SplitMix64(seed)
    return [rng.fr() for _ in range(n)]


# ------------------------------------------------------------------ primitives
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 255, 2049, 5000])
def test_poly_eval_batch(ctx, n):
    polys = [rand_vec(10 * n + i, max(1, n - i)) for i in range(5)]
    bufs = [ctx.upload(fr_to_mont(p)) for p in polys]
    pt = rand_vec(77, 1)[0]
    got = fr_from_mont(ctx.poly_eval([ctx.ref(b) for b in bufs], fr_to_mont1(pt)))
    assert got == [poly_eval(p, pt) for p in polys]


def test_poly_eval_large_and_offsets(ctx):
    n = (1 << 17) + 3
    p = rand_vec(5, n)
    buf = ctx.upload(fr_to_mont(p))
    pt = rand_vec(6, 1)[0]
    got = fr_from_mont(ctx.poly_eval([ctx.ref(buf), ctx.ref(buf, 100, 5000), ctx.ref(buf, n - 1, 1)], fr_to_mont1(pt)))
    assert got == [poly_eval(p, pt), poly_eval(p[100:5100], pt), p[-1]]


def test_poly_eval_two_points_one_launch(ctx):
    """zkp_poly_eval2_dev: 25 ragged polynomials (the shape of a proof's openings), each at one of two points."""
    lens = [1 << 15] + [4099] * 4 + [4096] * 16 + [4099, 4098, 1, 8193]
    polys = [rand_vec(900 + i, l) for i, l in enumerate(lens)]
    bufs = [ctx.upload(fr_to_mont(p)) for p in polys]
    z0, z1 = rand_vec(55, 2)
    which = [0] * 21 + [1] * 4
    got = fr_from_mont(ctx.poly_eval2([ctx.ref(b) for b in bufs], which, np.stack([fr_to_mont1(z0), fr_to_mont1(z1)])))
    assert got == [poly_eval(p, z1 if w else z0) for p, w in zip(polys, which)]


@pytest.mark.parametrize("n", [2, 3, 16, 17, 33, 1000, 16385, 32768 + 5, 100003])
def test_div_linear_is_ruffini(ctx, n):
    p = rand_vec(n, n)
    pt = rand_vec(n + 1, 1)[0]
    src = ctx.upload(fr_to_mont(p))
    dst = ctx.alloc(n)
    ctx.poly_div_linear(ctx.ref(src), fr_to_mont1(pt), dst)
    assert fr_from_mont(dst.download(0, n - 1)) == oplonk.ruffini(p, pt)


def test_lincomb_ragged(ctx):
    lens = [7, 100, 64, 1]
    polys = [rand_vec(20 + i, l) for i, l in enumerate(lens)]
    sc = rand_vec(30, 4)
    bufs = [ctx.upload(fr_to_mont(p)) for p in polys]
    out = ctx.alloc(120)
    ctx.poly_lincomb([ctx.ref(b) for b in bufs], fr_to_mont(sc), out, 10, 110)
    exp = [sum(s * (p[i] if i < len(p) else 0) for s, p in zip(sc, polys)) % R_MOD for i in range(110)]
    assert fr_from_mont(out.download(10, 110)) == exp


def test_blind_and_fill(ctx):
    n = 64
    p = rand_vec(1, n)
    for cnt in (2, 3):
        bl = rand_vec(2 + cnt, cnt)
        buf = ctx.alloc(n + cnt)
        buf.upload(fr_to_mont(p))
        ctx.poly_blind(buf, 0, n, fr_to_mont(bl))
        assert fr_from_mont(buf.download()) == oplonk.blind(p, bl, n)
    v = rand_vec(9, 1)[0]
    buf = ctx.alloc(100)
    buf.zero()
    ctx.fill(buf, 3, 90, fr_to_mont1(v))
    assert fr_from_mont(buf.download()) == [0] * 3 + [v] * 90 + [0] * 7


@pytest.mark.parametrize("k", [0, 1, 4, 5, 9, 12])
def test_perm_z_matches_reference_loop(ctx, k):
    """src/permutation.rs:205-300 with per-gate inversions vs the scan formulation."""
    n = 1 << k
    fft = OFft(k)
    wires = [rand_vec(40 + j, n) for j in range(4)]
    sig = [rand_vec(50 + j, n) for j in range(4)]
    beta, gamma = rand_vec(60, 2)
    roots = ctx.fft_elements(k)
    wb = [ctx.upload(fr_to_mont(w)) for w in wires]
    sb = [ctx.upload(fr_to_mont(s)) for s in sig]
    out = ctx.alloc(n)
    ctx.perm_z(n, [ctx.ref(b) for b in wb], [ctx.ref(b) for b in sb], roots, fr_to_mont1(beta), fr_to_mont1(gamma), out)
    assert fr_from_mont(out.download()) == oplonk.compute_permutation_vec(fft, wires, beta, gamma, sig)


def test_perm_lagrange(ctx):
    k = 6
    n = 1 << k
    rng = np.random.default_rng(1)
    w = rng.integers(0, 4, n)
    g = rng.integers(0, n, n)
    enc = (w.astype(np.uint32) << np.uint32(30)) | g.astype(np.uint32)
    roots = ctx.fft_elements(k)
    out = ctx.alloc(n)
    ctx.perm_lagrange(k, enc, roots, out)
    om = domain_generator(k)
    ks = (1, K1, K2, K3)
    assert fr_from_mont(out.download()) == [ks[int(a)] * pow(om, int(b), R_MOD) % R_MOD for a, b in zip(w, g)]


# ------------------------------------------------------------------ full proofs
CIRCUITS = {
    "range": lambda: circuits.range_circuit((1 << 64) - 1),
    "logic_curve": circuits.logic_curve_circuit,
    "readme": circuits.readme_circuit,
    "chain": lambda: circuits.arithmetic_chain(1000),
    "boolean_select": circuits.boolean_select_circuit,
    "decomposition": circuits.decomposition_circuit,
}


def both_sides(ctx, cs, label=b"demo", seed=8349):
    circ = SynthesizedCircuit.from_composer(cs)
    rng = SplitMix64(seed)
    tau = rng.fr()
    k = circ.n.bit_length() - 1
    pp = PlonkParams.setup_synthetic(ctx, max(k, 4) + 1, fr_to_mont1(tau))
    prover = z.PlonkKey.compile_with_circuit(pp, label, circ)
    commit = oplonk.default_commit(tau=tau)
    opk, ovk = oplonk.compile_circuit(circ, commit, pp.srs.n)
    otr = OTranscript.base(label, oplonk.vk_transcript_list(ovk), circ.m)   # the oracle hashes with its own Merlin
    bl = [rng.fr() for _ in range(11)]
    return circ, tau, prover, commit, opk, ovk, otr, bl


@pytest.mark.parametrize("name", list(CIRCUITS))
def test_compile_and_prove_bit_exact(ctx, name):
    circ, tau, prover, commit, opk, ovk, otr, bl = both_sides(ctx, CIRCUITS[name]())
    # key preprocessing: 15 commitments and every resident polynomial (src/key.rs)
    for nm in list(oplonk.SELECTORS) + ["s_sigma_%d" % i for i in (1, 2, 3, 4)]:
        assert prover.verifier_key[nm] == ovk[nm], nm
        assert fr_from_mont(prover.prover_key.poly[nm].download()) == opk.poly[nm], nm
    for nm in ("q_m", "q_arith", "s_sigma_3", "linear"):
        assert fr_from_mont(prover.prover_key.eval8[nm].download()) == opk.eval8[nm], nm
    otrace, gtrace = {}, {}
    oproof, opi = oplonk.create_proof(opk, circ, commit, otr, bl, trace=otrace)
    gproof, gpi = prover.create_proof(bl, circ, trace=gtrace)
    assert gpi == opi
    assert gtrace["challenges"] == otrace["challenges"] and gtrace["z_challenge"] == otrace["z_challenge"]
    ws = gtrace["workspace"]
    n = circ.n
    for j in range(4):
        assert fr_from_mont(ws["wp"][j].download()) == otrace["w_polys"][j]
    assert fr_from_mont(ws["Z"].download()) == otrace["z_evals"]
    assert fr_from_mont(ws["zp"].download()) == otrace["z_poly"]
    assert fr_from_mont(ws["T"].download()) == otrace["t_poly"]
    r = otrace["r_poly"]
    assert fr_from_mont(ws["R"].download()) == r + [0] * (n + 3 - len(r))
    wz = otrace["w_z_poly"]
    assert fr_from_mont(ws["WZ"].download(0, len(wz))) == wz
    assert fr_from_mont(ws["WZW"].download(0, n + 2)) == otrace["w_zw_poly"]
    assert gtrace["t_eval"] == otrace["t_eval"]
    for c in oplonk.Proof.COMM_NAMES:
        assert getattr(gproof, c) == getattr(oproof, c), c
    assert gproof.evaluations == oproof.evaluations
    # and the restated verifier accepts the GPU proof
    assert oplonk.verify(ovk, n, gproof, circ.pi_indexes, gpi, otr, oplonk.trapdoor_kzg_check(tau))
    if name == "readme":   # once with the real two-pairing batch_check (src/commitment_scheme.rs:52-64)
        from oracle import pairing
        assert oplonk.verify(ovk, n, gproof, circ.pi_indexes, gpi, otr,
                             pairing.kzg_pairing_check(pairing.g2_mul(pairing.G2_GEN, tau)))


def test_unsatisfied_circuit_errs(ctx):
    """tests/range.rs:79-85: the only Err path is commit's degree check."""
    circ, tau, prover, commit, opk, ovk, otr, bl = both_sides(ctx, circuits.range_circuit(7))
    bad = SynthesizedCircuit.from_composer(circuits.range_circuit((-(1 << 77)) % R_MOD))
    with pytest.raises(Error):
        prover.create_proof(bl, bad)
    with pytest.raises(oplonk.ProverError):
        oplonk.create_proof(opk, bad, commit, otr, bl)
    # the prover object is still usable afterwards
    gproof, gpi = prover.create_proof(bl, circ)
    assert oplonk.verify(ovk, circ.n, gproof, circ.pi_indexes, gpi, otr, oplonk.trapdoor_kzg_check(tau))


def test_prover_is_deterministic_and_blinders_matter(ctx):
    circ, tau, prover, commit, opk, ovk, otr, bl = both_sides(ctx, circuits.range_circuit(123456))
    p1, _ = prover.create_proof(bl, circ)
    p2, _ = prover.create_proof(bl, circ)
    assert p1 == p2
    p3, pi = prover.create_proof([b + 1 for b in bl], circ)
    assert p3.a_comm != p1.a_comm
    assert oplonk.verify(ovk, circ.n, p3, circ.pi_indexes, pi, otr, oplonk.trapdoor_kzg_check(tau))


def test_prove_2p16_gates_bit_exact_vs_c_prover(ctx, cport):
    """BASELINE config 2: the synthetic 2^16-gate circuit.  Every commitment of the key and the whole
    proof (11 commitments + 16 evaluations) equal the threaded C restatement; the restated verifier
    accepts the GPU proof."""
    from host_mirror.composer import synthetic_circuit
    from oracle import cprover, curve
    from oracle.fields import fr_to_raw_limbs, g1_to_mont_limbs
    k = 16
    circ = synthetic_circuit(k)
    rng = SplitMix64(8349)
    tau = rng.fr()
    pp = PlonkParams.setup_synthetic(ctx, k, fr_to_mont1(tau))
    prover = z.PlonkKey.compile(pp, circ)
    # the CPU side gets the very same SRS points the GPU generated
    cp = cprover.CProver(circ, pp.srs.download(), b"plonk", OTranscript)
    for nm in list(oplonk.SELECTORS) + ["s_sigma_%d" % i for i in (1, 2, 3, 4)]:
        assert prover.verifier_key[nm] == cp.vk[nm], nm
    bl = [rng.fr() for _ in range(11)]
    ctrace = {}
    cproof, cpi = cp.create_proof(bl, circ, trace=ctrace)
    gtrace = {}
    gproof, gpi = prover.create_proof(bl, circ, trace=gtrace)
    assert gpi == cpi
    assert np.array_equal(gtrace["workspace"]["T"].download(), ctrace["t_poly"])
    for c in oplonk.Proof.COMM_NAMES:
        assert getattr(gproof, c) == getattr(cproof, c), c
    assert gproof.evaluations == cproof.evaluations
    vk = dict(cp.vk)
    tr = OTranscript.base(b"plonk", oplonk.vk_transcript_list(vk), circ.m)
    assert oplonk.verify(vk, circ.n, gproof, circ.pi_indexes, gpi, tr, oplonk.trapdoor_kzg_check(tau))


def test_mul_point_circuit_bit_exact_vs_c_prover(ctx, cport):
    """tests/ecc.rs mul_point: 2025 gates of boolean / arithmetic / curve-addition rows (n = 2^11)."""
    from oracle import cprover
    circ = SynthesizedCircuit.from_composer(circuits.mul_point_circuit())
    rng = SplitMix64(8349)
    tau = rng.fr()
    pp = PlonkParams.setup_synthetic(ctx, 12, fr_to_mont1(tau))
    prover = z.PlonkKey.compile_with_circuit(pp, b"demo", circ)
    cp = cprover.CProver(circ, prover.keypair.srs.download(), b"demo", OTranscript)
    for nm in list(oplonk.SELECTORS) + ["s_sigma_%d" % i for i in (1, 2, 3, 4)]:
        assert prover.verifier_key[nm] == cp.vk[nm], nm
    bl = [rng.fr() for _ in range(11)]
    cproof, cpi = cp.create_proof(bl, circ)
    gproof, gpi = prover.create_proof(bl, circ)
    assert gpi == cpi and len(gpi) == 2
    for c in oplonk.Proof.COMM_NAMES:
        assert getattr(gproof, c) == getattr(cproof, c), c
    assert gproof.evaluations == cproof.evaluations
    vk = dict(cp.vk)
    tr = OTranscript.base(b"demo", oplonk.vk_transcript_list(vk), circ.m)
    assert oplonk.verify(vk, circ.n, gproof, circ.pi_indexes, gpi, tr, oplonk.trapdoor_kzg_check(tau))
    assert z.Proof.from_bytes(gproof.to_bytes()) == gproof


@pytest.mark.parametrize("name", ["logic", "boolean", "decomposition_bit", "public_input"])
def test_negative_cases_of_the_reference_suites(ctx, name):
    """tests/logic.rs:109-111, tests/boolean.rs:90, tests/decomposition.rs:101-103: a prover compiled for the
    honest circuit errs on an unsatisfying witness of the same shape; a wrong public input makes the
    verifier reject."""
    if name == "logic":
        good, bad = circuits.logic_curve_circuit(), circuits.logic_curve_circuit()
        # same gates, wrong product quad in the first logic row (tests/logic.rs negative case)
        row = next(c for c in bad.constraints if c.q_logic)
        bad.witness[row.w_o] = (bad.witness[row.w_o] + 1) % R_MOD
    elif name == "boolean":
        good, bad = circuits.boolean_select_circuit(bit=1), circuits.boolean_select_circuit(bit=2)
    elif name == "decomposition_bit":
        good = circuits.decomposition_circuit(23, 64)
        bad = circuits.decomposition_circuit(23, 64)
        # corrupt one bit witness after synthesis: same gates, unsatisfied boolean + sum rows
        bit_w = bad.constraints[6].w_a
        bad.witness[bit_w] = 2
    else:
        good = bad = circuits.range_circuit(99)
    circ, tau, prover, commit, opk, ovk, otr, bl = both_sides(ctx, good)
    if name == "public_input":
        good2 = circuits.readme_circuit()
        circ, tau, prover, commit, opk, ovk, otr, bl = both_sides(ctx, good2)
        proof, pi = prover.create_proof(bl, circ)
        assert oplonk.verify(ovk, circ.n, proof, circ.pi_indexes, pi, otr, oplonk.trapdoor_kzg_check(tau))
        wrong = [(pi[0] + 1) % R_MOD] + pi[1:]
        with pytest.raises(oplonk.VerifyError):
            oplonk.verify(ovk, circ.n, proof, circ.pi_indexes, wrong, otr, oplonk.trapdoor_kzg_check(tau))
        return
    badc = SynthesizedCircuit.from_composer(bad)
    assert badc.m == circ.m
    with pytest.raises(Error):
        prover.create_proof(bl, badc)
    with pytest.raises(oplonk.ProverError):
        oplonk.create_proof(opk, badc, commit, otr, bl)


@pytest.mark.parametrize("name", ["mul_generator", "add_point", "mul_point"])
def test_negative_ecc_cases_of_the_reference(ctx, name):
    """tests/ecc.rs:81-97 (mul_generator: b = 8 G for a = 7), :216-232 (add_point: c = 9 G != 7 G + 8 G),
    :303-318 (mul_point: unrelated c): the honest key proves and verifies the honest witness, and
    ``create_proof`` returns Err for a witness of the same shape that violates the curve relation -- on the
    GPU and in the oracle alike."""
    from host_mirror.composer import JUBJUB_GENERATOR as G, jubjub_mul
    if name == "mul_generator":
        good, bad = circuits.ecc_mul_generator_circuit(7), circuits.ecc_mul_generator_circuit(7, jubjub_mul(G, 8))
    elif name == "add_point":
        good, bad = circuits.ecc_add_point_circuit(), circuits.ecc_add_point_circuit(c=jubjub_mul(G, 9))
        ident = circuits.ecc_add_point_circuit(a=jubjub_mul(G, 5), b=(0, 1), c=jubjub_mul(G, 5))   # :175-194 identity works
    else:
        good, bad = circuits.ecc_mul_point_circuit(7), circuits.ecc_mul_point_circuit(7, c=jubjub_mul(G, 12345))
    circ, tau, prover, commit, opk, ovk, otr, bl = both_sides(ctx, good)
    proof, pi = prover.create_proof(bl, circ)
    oproof, _ = oplonk.create_proof(opk, circ, commit, otr, bl)
    assert proof.to_bytes() == oproof.to_bytes()
    assert oplonk.verify(ovk, circ.n, proof, circ.pi_indexes, pi, otr, oplonk.trapdoor_kzg_check(tau))
    if name == "add_point":
        ic = SynthesizedCircuit.from_composer(ident)
        assert ic.m == circ.m
        iproof, ipi = prover.create_proof(bl, ic)
        assert oplonk.verify(ovk, circ.n, iproof, circ.pi_indexes, ipi, otr, oplonk.trapdoor_kzg_check(tau))
    badc = SynthesizedCircuit.from_composer(bad)
    assert badc.m == circ.m
    with pytest.raises(Error):
        prover.create_proof(bl, badc)
    with pytest.raises(oplonk.ProverError):
        oplonk.create_proof(opk, badc, commit, otr, bl)


def _synthetic_setup(ctx, k):
    from host_mirror.composer import synthetic_circuit
    circ = synthetic_circuit(k)
    rng = SplitMix64(8349)
    tau = rng.fr()
    pp = PlonkParams.setup_synthetic(ctx, k, fr_to_mont1(tau))
    prover = z.PlonkKey.compile(pp, circ)
    bl = [rng.fr() for _ in range(11)]
    return circ, tau, pp, prover, bl


@pytest.mark.parametrize("name", ["range", "readme", "logic", "synthetic12"])
def test_compile_pair_prove_verify_on_the_product_side(ctx, name):
    """The reference's integration flow with no oracle in the loop (tests/range.rs:24-97, README):
    ``PublicParameters::setup -> PlonkKey::compile -> (prover, verifier)``, ``create_proof`` on the GPU,
    ``verifier.verify`` in the library's host code (csrc/verifier.cu: real pairing); wrong public inputs and a
    foreign proof are rejected.  The oracle's verifier agrees on every decision."""
    from host_mirror.composer import synthetic_circuit
    rng = SplitMix64(31)
    tau = rng.fr()
    if name == "synthetic12":
        circ, label = synthetic_circuit(12), b"plonk"
    else:
        comp = {"range": lambda: circuits.range_circuit(77), "readme": circuits.readme_circuit,
                "logic": circuits.logic_curve_circuit}[name]()
        circ, label = SynthesizedCircuit.from_composer(comp), b"demo"
    k = circ.n.bit_length() - 1
    pp = PlonkParams.setup_synthetic(ctx, k + 1, fr_to_mont1(tau))
    prover, verifier = z.PlonkKey.compile_pair(pp, circ, label)
    bl = [rng.fr() for _ in range(11)]
    proof, pi = prover.create_proof(bl, circ)
    verifier.verify(proof, pi)
    otr = OTranscript.base(label, oplonk.vk_transcript_list(prover.verifier_key), circ.m)
    assert oplonk.verify(prover.verifier_key, circ.n, proof, circ.pi_indexes, pi, otr, oplonk.trapdoor_kzg_check(tau))
    if pi:
        with pytest.raises(Error):
            verifier.verify(proof, [(pi[0] + 1) % R_MOD] + pi[1:])
    other, _ = prover.create_proof([b + 1 for b in bl], circ)
    verifier.verify(other, pi)                                  # a different blinding of the same statement
    forged = z.Proof.from_bytes(other.to_bytes())
    forged.w_z_chall_comm = proof.w_z_chall_w_comm
    with pytest.raises(Error):
        verifier.verify(forged, pi)
    prover.close()


@pytest.mark.parametrize("k", [12, 16, 20])
def test_synthetic_proof_equals_committed_oracle_digest(ctx, k):
    """The proofs bench.py times (seed 8349, BASELINE configs 2 and 5): wire bytes of the native driver against
    the sha256 the CPU oracle produced for the same circuit (tests/golden/make_synthetic_digests.py) -- no
    oracle code in this test."""
    import hashlib
    import json
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "synthetic_proofs.json")))
    circ, tau, pp, prover, bl = _synthetic_setup(ctx, k)
    proof, _ = prover.create_proof(bl, circ)
    assert proof.wire_bytes == proof.to_bytes()
    assert hashlib.sha256(proof.wire_bytes).hexdigest() == gold[str(k)]["sha256"]
    prover.close()


def test_prove_2p20_gates_bit_exact_vs_c_prover(ctx, cport):
    """BASELINE config 5 on one GPU: the synthetic 2^20-gate circuit through the native driver.  All 15 key
    commitments, the 11 proof commitments, the 16 evaluations and the 1040 wire bytes equal the threaded C
    restatement run here on the same SRS points; the restated verifier accepts."""
    from oracle import cprover
    circ, tau, pp, prover, bl = _synthetic_setup(ctx, 20)
    cp = cprover.CProver(circ, pp.srs.download(), b"plonk", OTranscript)
    for nm in list(oplonk.SELECTORS) + ["s_sigma_%d" % i for i in (1, 2, 3, 4)]:
        assert prover.verifier_key[nm] == cp.vk[nm], nm
    cproof, cpi = cp.create_proof(bl, circ)
    gproof, gpi = prover.create_proof(bl, circ)
    assert gpi == cpi
    for c in oplonk.Proof.COMM_NAMES:
        assert getattr(gproof, c) == getattr(cproof, c), c
    assert gproof.evaluations == cproof.evaluations
    assert gproof.wire_bytes == cproof.to_bytes()
    prover.verifier().verify(gproof, gpi)       # product-side verifier (host pairing) accepts
    vk = dict(cp.vk)
    tr = OTranscript.base(b"plonk", oplonk.vk_transcript_list(vk), circ.m)
    assert oplonk.verify(vk, circ.n, gproof, circ.pi_indexes, gpi, tr, oplonk.trapdoor_kzg_check(tau))
    prover.close()


@pytest.mark.parametrize("name", ["range", "readme", "logic"])
def test_native_round_driver_equals_python_driven_rounds(ctx, name):
    """csrc/create_proof.cu (rounds + transcript + linearisation scalars in native code, one C-ABI call)
    against the same rounds driven call by call from Python -- which the tests above pin to the oracle:
    identical proofs, identical 1040-byte wire format, from host-resident and device-resident witnesses."""
    comp = {"range": lambda: circuits.range_circuit(424242), "readme": circuits.readme_circuit,
            "logic": circuits.logic_curve_circuit}[name]()
    circ, tau, prover, commit, opk, ovk, otr, bl = both_sides(ctx, comp)
    assert prover.native
    nproof, npi = prover.create_proof(bl, circ)
    assert nproof.wire_bytes == nproof.to_bytes()          # native serialisation == host mirror's
    prover.native = False
    pproof, ppi = prover.create_proof(bl, circ)
    prover.native = True
    assert nproof == pproof and npi == ppi
    oproof, opi = oplonk.create_proof(opk, circ, commit, otr, bl)
    for c in oplonk.Proof.COMM_NAMES:
        assert getattr(nproof, c) == getattr(oproof, c), c
    assert nproof.evaluations == oproof.evaluations
    # the three witness forms: values + device gather (above), gathered on the host, already in HBM
    hproof, _ = prover.create_proof(bl, z.WitnessAssignment.from_circuit(circ, circ.n))
    assert hproof == nproof and hproof.wire_bytes == nproof.wire_bytes
    wa = z.WitnessAssignment.from_circuit(circ, circ.n).to_device(ctx)
    dproof, _ = prover.create_proof(bl, wa)
    assert dproof == nproof and dproof.wire_bytes == nproof.wire_bytes
    vproof, _ = prover.create_proof(bl, z.WitnessValues.from_circuit(circ))
    assert vproof == nproof
    assert z.Proof.from_bytes(nproof.wire_bytes) == nproof
    assert oplonk.verify(ovk, circ.n, nproof, circ.pi_indexes, npi, otr, oplonk.trapdoor_kzg_check(tau))


def test_cloned_provers_prove_concurrently(ctx):
    """``Prover: Clone`` (src/prover.rs:28): three provers over one device-resident key, each on its own context and
    host thread, 6 proofs each with different blinders -- every proof equals the one the original prover
    makes alone for the same blinders (no cross-talk between streams / scratch), and verifies."""
    import threading
    circ, tau, pp, prover, bl = _synthetic_setup(ctx, 12)
    sets = [[(b + 17 * j) % R_MOD for b in bl] for j in range(6)]
    want = [prover.create_proof(b, circ)[0].wire_bytes for b in sets]
    clones = [prover, prover.clone(), prover.clone()]
    got = [[None] * len(sets) for _ in clones]
    errs = []

    def work(p, out):
        try:
            for j, b in enumerate(sets):
                out[j] = p.create_proof(b, circ)[0].wire_bytes
        except Exception as e:   # surfaced below: a thread must not die silently
            errs.append(e)
    th = [threading.Thread(target=work, args=(p, o)) for p, o in zip(clones, got)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for o in got:
        assert o == want
    ver = prover.verifier()
    ver.verify(z.Proof.from_bytes(got[2][5]), list(circ.pi_values))
    for p in clones[1:]:
        p.close()
    prover.close()

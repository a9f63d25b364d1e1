"""CPU: the restated prover / verifier (oracle/plonk.py), the host composer and the Merlin
transcript.  Follows the reference's integration tests (tests/range.rs:66-97 etc.):
positive -> create_proof + verify succeed; negative -> create_proof errs."""
import pytest

from host_mirror.composer import SELECTORS, SynthesizedCircuit, jubjub_on_curve, JUBJUB_GENERATOR
from dusk_plonk_b200.transcript import MerlinTranscript, Transcript
from dusk_plonk_b200 import widgets
from oracle import plonk, curve
from oracle.fields import R_MOD
from oracle.rng import SplitMix64

import circuits


def test_merlin_known_answer():
    """merlin crate's published conformance vector."""
    t = MerlinTranscript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == \
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_native_keccak_equals_python_spec():
    from dusk_plonk_b200.transcript import keccak_f1600, keccak_f1600_py
    import hashlib
    a = bytearray(hashlib.sha256(b"seed").digest() * 7)[:200]
    b = bytearray(a)
    for _ in range(3):
        keccak_f1600(a)
        keccak_f1600_py(b)
        assert a == b
    st = bytearray(200)
    st[0] ^= 0x06
    st[135] ^= 0x80
    keccak_f1600(st)
    assert st[:32].hex() == hashlib.sha3_256(b"").hexdigest()


def test_gate_counts_match_reference():
    """m = 18 for tests/range.rs (n = 32) and 287 for the README circuit."""
    assert jubjub_on_curve(JUBJUB_GENERATOR)
    assert circuits.range_circuit(7).m() == 18
    assert circuits.readme_circuit().m() == 287


def rows_violated(cs):
    sel = cs.selector_columns()
    W = cs.wire_indices()
    wv = [[cs.witness[i] for i in W[j]] + [0] for j in range(4)]
    bad = []
    for i in range(cs.m()):
        q = {s: sel[s][i] for s in SELECTORS}
        t = plonk.gate_identity(q, (5, 7, 11, 13), wv[0][i], wv[1][i], wv[2][i], wv[3][i],
                                wv[0][i + 1], wv[1][i + 1], wv[3][i + 1])
        if (t + cs.instance.get(i, 0)) % R_MOD:
            bad.append(i)
    return bad


@pytest.mark.parametrize("name", ["range", "readme", "logic_curve", "chain", "boolean_select", "decomposition",
                                  "mul_point"])
def test_widget_identities_vanish_on_honest_rows(name):
    cs = {"range": lambda: circuits.range_circuit((1 << 64) - 1), "readme": circuits.readme_circuit,
          "logic_curve": circuits.logic_curve_circuit, "chain": lambda: circuits.arithmetic_chain(64),
          "boolean_select": circuits.boolean_select_circuit, "decomposition": circuits.decomposition_circuit,
          "mul_point": circuits.mul_point_circuit}[name]()
    assert rows_violated(cs) == []


def test_gadget_gate_counts_match_reference_docs():
    """component_decomposition consumes 2 N + 1 gates (src/lib.rs:879); boolean is one gate."""
    base = circuits.range_circuit(1, 8).m() - 2 - 1      # initialize() alone: range(8 bits) adds 2 + assert 1
    assert base == 6
    assert circuits.decomposition_circuit(23, 64).m() == 6 + (2 * 64 + 1) + 64


def test_proof_wire_format_roundtrip():
    from dusk_plonk_b200.prover import Proof, COMM_NAMES
    circ, tau, commit, pk, vk, tr, bl = setup(circuits.range_circuit(12345))
    op, _ = plonk.create_proof(pk, circ, commit, tr, bl)
    p = Proof()
    for c in COMM_NAMES:
        setattr(p, c, getattr(op, c))
    p.evaluations = dict(op.evaluations)
    raw = p.to_bytes()
    assert len(raw) == 11 * 48 + 16 * 32
    q = Proof.from_bytes(raw)
    assert q == p and q.to_bytes() == raw
    ident = Proof.from_bytes(bytes([0xC0]) + bytes(47) + raw[48:])
    assert ident.a_comm is None


def test_widget_identities_catch_bad_witness():
    assert rows_violated(circuits.range_circuit((-(1 << 77)) % R_MOD)) != []
    assert rows_violated(circuits.logic_curve_circuit(bad=True)) != []


def setup(cs, label=b"demo", seed=8349):
    circ = SynthesizedCircuit.from_composer(cs)
    rng = SplitMix64(seed)
    tau = rng.fr()
    commit = plonk.default_commit(tau=tau)
    pk, vk = plonk.compile_circuit(circ, commit, 1 << 20)
    from oracle.merlin import Transcript as OTranscript
    tr = OTranscript.base(label, plonk.vk_transcript_list(vk), circ.m)
    bl = [rng.fr() for _ in range(11)]
    return circ, tau, commit, pk, vk, tr, bl


@pytest.mark.parametrize("name", ["range", "logic_curve", "readme", "boolean_select", "decomposition"])
def test_prove_and_verify(name):
    cs = {"range": lambda: circuits.range_circuit((1 << 64) - 1), "readme": circuits.readme_circuit,
          "logic_curve": circuits.logic_curve_circuit, "boolean_select": circuits.boolean_select_circuit,
          "decomposition": circuits.decomposition_circuit}[name]()
    circ, tau, commit, pk, vk, tr, bl = setup(cs)
    proof, pi = plonk.create_proof(pk, circ, commit, tr, bl)
    assert plonk.verify(vk, pk.n, proof, circ.pi_indexes, pi, tr, plonk.trapdoor_kzg_check(tau))
    # tampered evaluation / wrong public input are rejected
    good = proof.evaluations["a_eval"]
    proof.evaluations["a_eval"] = (good + 1) % R_MOD
    with pytest.raises(plonk.VerifyError):
        plonk.verify(vk, pk.n, proof, circ.pi_indexes, pi, tr, plonk.trapdoor_kzg_check(tau))
    proof.evaluations["a_eval"] = good
    if pi:
        with pytest.raises(plonk.VerifyError):
            plonk.verify(vk, pk.n, proof, circ.pi_indexes, [(pi[0] + 1) % R_MOD] + pi[1:], tr,
                         plonk.trapdoor_kzg_check(tau))


def test_negative_range_errs_like_reference():
    """tests/range.rs:79-85: a = -(2^77) is out of range -> create_proof returns Err."""
    circ, tau, commit, pk, vk, tr, bl = setup(circuits.range_circuit(7))
    bad = SynthesizedCircuit.from_composer(circuits.range_circuit((-(1 << 77)) % R_MOD))
    with pytest.raises(plonk.ProverError):
        plonk.create_proof(pk, bad, commit, tr, bl)


def test_trapdoor_commit_equals_pippenger():
    rng = SplitMix64(5)
    tau = rng.fr()
    pts = curve.srs_powers(tau, 40)
    coeffs = [rng.fr() for _ in range(37)] + [0, 0]
    a = plonk.default_commit(tau=tau)(coeffs, 40)
    b = plonk.default_commit(srs_points=pts)(coeffs, 40)
    assert a == b and a is not None


def test_host_widget_scalars_equal_oracle():
    rng = SplitMix64(99)
    ch = tuple(rng.fr() for _ in range(8))
    ev = {k: rng.fr() for k in plonk.EVAL_NAMES}
    assert widgets.linearization_scalars(512, ch, ev) == plonk.linearization_scalars(512, ch, ev)

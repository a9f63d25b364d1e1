"""CPU: the product-side verifier glue (csrc/verifier.cu, host code: no GPU needed) -- the BLS12-381 pairing
against bilinearity and the oracle's pairing, [tau]_2 against the oracle's G2 arithmetic, and
``Verifier::verify`` (src/verifier.rs:46-81, src/prover/proof.rs:70-383) on the committed golden proofs:
accepts them, rejects tampered proofs / wrong public inputs / a wrong opening key, as the reference's
integration tests require (tests/range.rs:66-97 and the negative cases)."""
import ctypes
import json
import os

import numpy as np
import pytest

import circuits
import dusk_plonk_b200 as z
from dusk_plonk_b200.field import P_MOD, R_MOD, fr_to_mont1, g1_from_mont, g1_to_mont
from dusk_plonk_b200.key import VerificationKey
from dusk_plonk_b200.plonk_params import Error
from dusk_plonk_b200.verifier import EvaluationKey, Verifier
from host_mirror.composer import SynthesizedCircuit
from oracle import curve, pairing
from oracle.rng import SplitMix64

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.json")))
_R384 = (1 << 384) % P_MOD


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def g2_to_limbs(pt):
    if pt is None:
        return np.zeros(24, dtype=np.uint64)
    (x0, x1), (y0, y1) = pt
    raw = b"".join((c * _R384 % P_MOD).to_bytes(48, "little") for c in (x0, x1, y0, y1))
    return np.frombuffer(raw, dtype="<u8").copy()


def g2_from_limbs(w):
    raw = np.ascontiguousarray(w, dtype="<u8").tobytes()
    rinv = pow(_R384, -1, P_MOD)
    c = [int.from_bytes(raw[48 * i:48 * (i + 1)], "little") * rinv % P_MOD for i in range(4)]
    return None if not any(c) else ((c[0], c[1]), (c[2], c[3]))


@pytest.fixture(scope="module")
def lib():
    return z.load_library()


def test_generator_multiples_match_oracle(lib):
    rng = SplitMix64(3)
    for s in (1, 2, R_MOD - 1, rng.fr(), rng.fr()):
        out2 = np.zeros(24, dtype=np.uint64)
        assert lib.zkp_g2_generator_mul(_ptr(fr_to_mont1(s)), _ptr(out2)) == 0
        assert g2_from_limbs(out2) == pairing.g2_mul(pairing.G2_GEN, s)
        assert pairing.g2_on_curve(g2_from_limbs(out2))
        out1 = np.zeros(12, dtype=np.uint64)
        assert lib.zkp_g1_generator_mul(_ptr(fr_to_mont1(s)), _ptr(out1)) == 0
        assert g1_from_mont(out1) == curve.mul(curve.G1_GEN, s)
    out2 = np.zeros(24, dtype=np.uint64)
    assert lib.zkp_g2_generator_mul(_ptr(fr_to_mont1(0)), _ptr(out2)) == 0 and not out2.any()


def test_pairing_is_bilinear_and_agrees_with_the_oracle(lib):
    rng = SplitMix64(9)
    a, b = rng.fr(), rng.fr()
    def g1(s): return g1_to_mont(curve.mul(curve.G1_GEN, s % R_MOD))
    def g2(s): return g2_to_limbs(pairing.g2_mul(pairing.G2_GEN, s % R_MOD))
    def check(pairs):
        p = np.ascontiguousarray(np.stack([x for x, _ in pairs]))
        q = np.ascontiguousarray(np.stack([y for _, y in pairs]))
        return lib.zkp_pairing_check(_ptr(p), _ptr(q), len(pairs))
    # e(aG, bH) e(-abG, H) = 1
    assert check([(g1(a), g2(b)), (g1(-a * b), g2(1))]) == 0
    assert check([(g1(a), g2(b)), (g1(-a * b + 1), g2(1))]) == z.ZKP_ERR_VERIFY
    # e(aG, H) e(G, -aH) = 1; three factors; the identity contributes 1
    assert check([(g1(a), g2(1)), (g1(1), g2(-a))]) == 0
    assert check([(g1(a), g2(1)), (g1(b), g2(1)), (g1(1), g2(-(a + b)))]) == 0
    assert check([(np.zeros(12, dtype=np.uint64), g2(5)), (g1(7), np.zeros(24, dtype=np.uint64))]) == 0
    assert check([(g1(1), g2(1))]) == z.ZKP_ERR_VERIFY           # non-degenerate
    # same accept / reject decisions as the oracle's (independent, polynomial-basis) pairing
    P1, Q1 = curve.mul(curve.G1_GEN, a), pairing.g2_mul(pairing.G2_GEN, b)
    P2, Q2 = curve.mul(curve.G1_GEN, (-a * b) % R_MOD), pairing.G2_GEN
    assert pairing.pairing_product_is_one([(P1, Q1), (P2, Q2)])
    bad = np.zeros(24, dtype=np.uint64); bad[0] = 5
    assert check([(g1(1), bad)]) == z.ZKP_ERR_INVALID            # not on the twist


def _golden_case(entry):
    build = {"range": lambda: circuits.range_circuit((1 << 64) - 1), "readme": circuits.readme_circuit,
             "logic_curve": circuits.logic_curve_circuit}[entry["circuit"]]
    circ = SynthesizedCircuit.from_composer(build())
    vk = VerificationKey()
    for nm, pt in entry["verifier_key"].items():
        vk[nm] = None if pt is None else (int(pt[0], 16), int(pt[1], 16))
    vk["n"] = circ.m
    proof = z.Proof.from_bytes(bytes.fromhex(entry["proof_bytes"]))
    pis = [int(p, 16) for p in entry["public_inputs"]]
    ok = EvaluationKey.from_tau(fr_to_mont1(int(entry["tau"], 16)))
    return circ, vk, proof, pis, ok


@pytest.mark.parametrize("entry", GOLD["proofs"], ids=lambda e: e["circuit"])
def test_verifier_accepts_golden_proofs_and_rejects_tampering(entry):
    circ, vk, proof, pis, ok = _golden_case(entry)
    ver = Verifier(entry["label"].encode(), vk, ok, circ.pi_indexes, circ.n, circ.m)
    ver.verify(proof, pis)                                       # Ok(())
    # a changed evaluation, a changed commitment, a wrong public input, a wrong opening key
    t = z.Proof.from_bytes(proof.to_bytes())
    t.evaluations["a_eval"] = (t.evaluations["a_eval"] + 1) % R_MOD
    with pytest.raises(Error):
        ver.verify(t, pis)
    t = z.Proof.from_bytes(proof.to_bytes())
    t.z_comm = curve.add(t.z_comm, curve.G1_GEN)
    with pytest.raises(Error):
        ver.verify(t, pis)
    if pis:
        with pytest.raises(Error):
            ver.verify(proof, [(pis[0] + 1) % R_MOD] + pis[1:])
        with pytest.raises(Error):
            ver.verify(proof, pis[:-1])                          # InconsistentPublicInputsLen
    wrong = Verifier(entry["label"].encode(), vk, EvaluationKey.from_tau(fr_to_mont1(12345)), circ.pi_indexes,
                     circ.n, circ.m)
    with pytest.raises(Error):
        wrong.verify(proof, pis)

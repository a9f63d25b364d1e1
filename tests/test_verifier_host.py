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
    # the Frobenius-based final exponentiation == the cube of the plain (p^12 - 1) / r power, on several pairings
    for s1, s2 in ((1, 1), (a, b), (R_MOD - 1, 3)):
        assert lib.zkp_pairing_selftest(_ptr(g1(s1)), _ptr(g2(s2))) == 0


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


def test_native_proof_decoder_validates_untrusted_bytes(lib):
    """zkp_proof_decode / zkp_g1_decompress (host code): the golden proofs decode to the values the Python decoder
    gives; every malformed encoding the zcash / dusk ``from_compressed`` rejects is rejected -- missing flag,
    x >= p, stray bits with the infinity flag, x off the curve, a curve point outside the prime-order subgroup,
    a non-canonical scalar (src/prover/proof.rs:77)."""
    from dusk_plonk_b200.field import g1_decompress as py_decompress, g1_mul
    from dusk_plonk_b200.prover import COMM_NAMES
    from dusk_plonk_b200.transcript import g1_compress
    for entry in GOLD["proofs"]:
        raw = bytes.fromhex(entry["proof_bytes"])
        p = z.Proof.from_bytes(raw)
        for i, c in enumerate(COMM_NAMES):
            assert getattr(p, c) == py_decompress(raw[48 * i:48 * (i + 1)])
        assert p.to_bytes() == raw
    raw = bytearray(bytes.fromhex(GOLD["proofs"][0]["proof_bytes"]))

    def rejected(buf):
        try:
            z.Proof.from_bytes(bytes(buf))
        except ValueError:
            return True
        return False
    bad = bytearray(raw); bad[0] &= 0x7F                      # compression flag cleared
    assert rejected(bad)
    bad = bytearray(raw); bad[0:48] = bytes([0xC0]) + bytes(46) + b"\x01"   # infinity with a stray bit
    assert rejected(bad)
    bad = bytearray(raw); bad[0:48] = bytes([0x9F]) + b"\xff" * 47           # x >= p
    assert rejected(bad)
    bad = bytearray(raw); bad[48 * 11 + 31] = 0xFF                           # scalar >= r
    assert rejected(bad)
    ok = bytearray(raw); ok[0:48] = bytes([0xC0]) + bytes(47)                # a well-formed identity is fine
    assert not rejected(ok)
    # x off the curve, and a curve point of the wrong order (G1 has cofactor ~2^126: almost every curve point)
    off = on_not_sub = None
    x = 5
    while off is None or on_not_sub is None:
        y2 = (pow(x, 3, P_MOD) + 4) % P_MOD
        y = pow(y2, (P_MOD + 1) // 4, P_MOD)
        if y * y % P_MOD != y2:
            off = off or x
        elif g1_mul((x, y), R_MOD) is not None:      # (unreduced scalar: [r] P for a point of another order)
            on_not_sub = on_not_sub or (x, y)
        x += 1
    bad = bytearray(raw); bad[0:48] = bytes([0x80 | (off >> 376)]) + off.to_bytes(48, "big")[1:]
    assert rejected(bad)
    bad = bytearray(raw); bad[0:48] = g1_compress(on_not_sub)
    assert rejected(bad)
    out = np.zeros(12, dtype=np.uint64)
    good = np.frombuffer(g1_compress(curve.mul(curve.G1_GEN, 12345)), dtype=np.uint8).copy()
    assert lib.zkp_g1_decompress(_ptr(good), _ptr(out)) == 0
    assert g1_from_mont(out) == curve.mul(curve.G1_GEN, 12345)
    notsub = np.frombuffer(g1_compress(on_not_sub), dtype=np.uint8).copy()
    assert lib.zkp_g1_decompress(_ptr(notsub), _ptr(out)) == z.ZKP_ERR_INVALID

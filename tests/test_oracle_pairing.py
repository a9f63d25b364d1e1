"""CPU: the Python pairing behind the restated verifier's batch_check (oracle/pairing.py)."""
import pytest

from oracle import curve, pairing as pr
from oracle.fields import R_MOD


def test_g2_generator_and_field():
    assert pr.g2_on_curve(pr.G2_GEN)
    assert pr.g2_mul(pr.G2_GEN, R_MOD - 1) == (pr.G2_GEN[0], pr.f2_sub((0, 0), pr.G2_GEN[1]))
    a = pr.f12([3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5, 8])
    assert pr.f12_mul(a, pr.f12_inv(a)) == pr.F12_ONE


def test_bilinearity_and_non_degeneracy():
    a, b = 0x1234567, 0x89ABCDEF0123
    e1 = pr.pairing(pr.G2_GEN, curve.G1_GEN)
    assert e1 != pr.F12_ONE
    assert pr.f12_pow(e1, R_MOD) == pr.F12_ONE
    e_ab = pr.pairing(pr.g2_mul(pr.G2_GEN, b), curve.mul(curve.G1_GEN, a))
    assert e_ab == pr.f12_pow(e1, a * b % R_MOD)
    # product form used by batch_check: e(-aG, bH) * e(abG, H) == 1
    assert pr.pairing_product_is_one([(curve.neg(curve.mul(curve.G1_GEN, a)), pr.g2_mul(pr.G2_GEN, b)),
                                      (curve.mul(curve.G1_GEN, a * b % R_MOD), pr.G2_GEN)])
    assert not pr.pairing_product_is_one([(curve.neg(curve.mul(curve.G1_GEN, a)), pr.g2_mul(pr.G2_GEN, b)),
                                          (curve.mul(curve.G1_GEN, a * b + 1), pr.G2_GEN)])


def test_verifier_with_real_pairing_accepts_and_rejects():
    """Proof::verify + batch_check end to end with the 2-pairing product (tests/range.rs:66-76)."""
    import circuits
    from host_mirror.composer import SynthesizedCircuit
    from oracle.merlin import Transcript
    from oracle import plonk
    from oracle.rng import SplitMix64
    circ = SynthesizedCircuit.from_composer(circuits.range_circuit((1 << 64) - 1))
    rng = SplitMix64(8349)
    tau = rng.fr()
    commit = plonk.default_commit(tau=tau)
    pk, vk = plonk.compile_circuit(circ, commit, 1 << 20)
    tr = Transcript.base(b"demo", plonk.vk_transcript_list(vk), circ.m)
    proof, pi = plonk.create_proof(pk, circ, commit, tr, [rng.fr() for _ in range(11)])
    check = pr.kzg_pairing_check(pr.g2_mul(pr.G2_GEN, tau))
    assert plonk.verify(vk, pk.n, proof, circ.pi_indexes, pi, tr, check)
    proof.evaluations["d_eval"] = (proof.evaluations["d_eval"] + 1) % R_MOD
    with pytest.raises(plonk.VerifyError):
        plonk.verify(vk, pk.n, proof, circ.pi_indexes, pi, tr, check)

"""The threaded C restatement (oracle/zkp_oracle.c, the CPU baseline) against the Python
big-integer oracle."""
import numpy as np
import pytest

from oracle import curve, ntt
from oracle.fields import (P_MOD, R_MOD, fq_from_mont_limbs, fq_to_mont_limbs, fr_from_mont_limbs,
                           fr_to_mont_limbs, fr_to_raw_limbs, g1_from_mont_limbs, g1_to_mont_limbs)
from oracle.rng import SplitMix64


def test_field_ops(cport):
    rng = SplitMix64(11)
    vals = [0, 1, R_MOD - 1] + [rng.fr() for _ in range(60)]
    for i, a in enumerate(vals):
        b = vals[(7 * i + 3) % len(vals)]
        al, bl = fr_to_mont_limbs([a])[0], fr_to_mont_limbs([b])[0]
        assert fr_from_mont_limbs(cport.fr_mul(al, bl))[0] == a * b % R_MOD
        assert fr_from_mont_limbs(cport.fr_add(al, bl))[0] == (a + b) % R_MOD
        assert fr_from_mont_limbs(cport.fr_sub(al, bl))[0] == (a - b) % R_MOD
        x = (a * b * 977 + i) % P_MOD
        y = (a * a + b) % P_MOD
        assert fq_from_mont_limbs(cport.fq_mul(fq_to_mont_limbs([x])[0], fq_to_mont_limbs([y])[0]))[0] == x * y % P_MOD


def test_ntt_all_modes(cport):
    rng = SplitMix64(12)
    for k in (0, 1, 2, 6, 9):
        f = ntt.Fft(k)
        v = [rng.fr() for _ in range(max(1, (1 << k) - 3))]
        vm = fr_to_mont_limbs(v)
        assert fr_from_mont_limbs(cport.ntt(vm, k)) == f.dft(v)
        assert fr_from_mont_limbs(cport.ntt(vm, k, inverse=True)) == f.idft(v)
        assert fr_from_mont_limbs(cport.ntt(vm, k, coset=True)) == f.coset_dft(v)
        assert fr_from_mont_limbs(cport.ntt(vm, k, inverse=True, coset=True)) == f.coset_idft(v)


def test_msm_and_fixed_base(cport):
    rng = SplitMix64(13)
    n = 200
    tau = rng.fr()
    dl = [pow(tau, i, R_MOD) for i in range(n)]
    g = g1_to_mont_limbs([curve.G1_GEN])[0]
    bases = cport.fixed_base_mul(g, fr_to_raw_limbs(dl))
    assert g1_from_mont_limbs(bases[:4]) == curve.srs_powers(tau, 4)
    sc = [rng.fr() for _ in range(n)]
    sc[3], sc[5], sc[7] = 0, 1, R_MOD - 1
    out = cport.msm_g1(bases, fr_to_mont_limbs(sc))
    assert g1_from_mont_limbs(out)[0] == curve.commit_known_dlog(dl, sc)
    # repeated bases force the doubling / cancellation branches; an infinity base is skipped
    b2 = np.tile(g, (50, 1))
    b2[10] = 0
    assert g1_from_mont_limbs(cport.msm_g1(b2, fr_to_mont_limbs([1] * 50)))[0] == curve.mul(curve.G1_GEN, 49)
    s3 = [1] * 25 + [R_MOD - 1] * 25
    b3 = np.tile(g, (50, 1))
    assert g1_from_mont_limbs(cport.msm_g1(b3, fr_to_mont_limbs(s3)))[0] is None
    assert g1_from_mont_limbs(cport.msm_g1(b3[:0], fr_to_mont_limbs([])))[0] is None


# ---------------------------------------------------------------- C prover vs Python prover
def _srs(cport, tau, n):
    from oracle import curve
    from oracle.fields import fr_to_raw_limbs, g1_to_mont_limbs
    dl, t = [], 1
    for _ in range(n):
        dl.append(t)
        t = t * tau % R_MOD
    return cport.fixed_base_mul(g1_to_mont_limbs([curve.G1_GEN])[0], fr_to_raw_limbs(dl))


@pytest.mark.parametrize("name,faithful", [("range", False), ("range", True), ("logic_curve", False), ("readme", False)])
def test_c_prover_equals_python_prover(cport, name, faithful):
    import circuits
    from host_mirror.composer import SynthesizedCircuit
    from oracle.merlin import Transcript
    from oracle import plonk, cprover
    from oracle.rng import SplitMix64
    cs = {"range": lambda: circuits.range_circuit((1 << 64) - 1), "readme": circuits.readme_circuit,
          "logic_curve": circuits.logic_curve_circuit}[name]()
    circ = SynthesizedCircuit.from_composer(cs)
    rng = SplitMix64(8349)
    tau = rng.fr()
    commit = plonk.default_commit(tau=tau)
    pk, vk = plonk.compile_circuit(circ, commit, 1 << 20)
    tr = Transcript.base(b"demo", plonk.vk_transcript_list(vk), circ.m)
    bl = [rng.fr() for _ in range(11)]
    t_py, t_c = {}, {}
    p_py, pi = plonk.create_proof(pk, circ, commit, tr, bl, trace=t_py)
    cp = cprover.CProver(circ, _srs(cport, tau, pk.max_len), b"demo", Transcript, faithful=faithful)
    for nm in plonk.SELECTORS:
        assert cp.vk[nm] == vk[nm], nm
    p_c, pi_c = cp.create_proof(bl, circ, trace=t_c)
    assert pi_c == pi
    assert fr_from_mont_limbs(t_c["t_poly"]) == t_py["t_poly"]
    assert fr_from_mont_limbs(t_c["z_evals"]) == t_py["z_evals"]
    assert fr_from_mont_limbs(t_c["w_z_poly"]) == t_py["w_z_poly"][:len(t_c["w_z_poly"])]
    assert p_c == p_py
    assert plonk.verify(vk, pk.n, p_c, circ.pi_indexes, pi_c, tr, plonk.trapdoor_kzg_check(tau))

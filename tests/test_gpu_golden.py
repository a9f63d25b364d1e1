"""GPU parity against the committed golden vectors (tests/golden/vectors.json) with NO oracle code
in the loop: inputs are re-derived from their seeds here, outputs compared with the stored values /
digests.  NTT kinds as Fft::{dft, idft, coset_dft, coset_idft}, commits as PlonkParams::commit, whole
proofs as the 1040-byte wire format of Prover::create_proof (src/prover/proof.rs:36-66)."""
import hashlib
import json
import os

import numpy as np
import pytest

import dusk_plonk_b200 as z
from host_mirror.composer import SynthesizedCircuit
from dusk_plonk_b200.field import R_MOD, fr_from_mont, fr_to_mont, fr_to_mont1, g1_from_mont
from dusk_plonk_b200.plonk_params import PlonkParams

import circuits

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vectors.json")))
M64 = (1 << 64) - 1


class SplitMix64:
    """The seeded generator the fixtures were drawn from (restated: no oracle import in this file)."""

    def __init__(self, seed):
        self.s = seed & M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & M64
        v = self.s
        v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & M64
        v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & M64
        return v ^ (v >> 31)

    def fr(self):
        while True:
            v = self.next() | (self.next() << 64) | (self.next() << 128) | ((self.next() >> 1) << 192)
            if v < R_MOD:
                return v


def digest(vals):
    h = hashlib.sha256()
    for x in vals:
        h.update(int(x).to_bytes(32, "little"))
    return h.hexdigest()


def check(entry, name, got):
    if name in entry:
        assert ["%064x" % x for x in got] == entry[name], name
    else:
        assert digest(got) == entry[name + "_sha256"], name


@pytest.mark.parametrize("entry", GOLD["ntt"], ids=lambda e: "k%d" % e["k"])
def test_ntt_kinds_match_golden(ctx, entry):
    k, rng = entry["k"], SplitMix64(entry["seed"])
    v = fr_to_mont([rng.fr() for _ in range(entry["len_in"])])
    if entry["len_in"] > (1 << k):      # blinded-size polynomial into the 8n coset (quotient_poly.rs:54-58)
        check(entry, "coset_dft_8n", fr_from_mont(ctx.ntt(v, k + 3, False, True)))
        return
    for name, inverse, coset in (("dft", False, False), ("idft", True, False), ("coset_dft", False, True),
                                 ("coset_idft", True, True)):
        check(entry, name, fr_from_mont(ctx.ntt(v, k, inverse, coset)))


@pytest.mark.parametrize("entry", GOLD["msm"], ids=lambda e: "n%d" % e["n"])
def test_commit_matches_golden(ctx, entry):
    n, rng = entry["n"], SplitMix64(entry["seed"])
    tau = rng.fr()
    assert "%064x" % tau == entry["tau"]
    sc = [rng.fr() for _ in range(n)]
    if n >= 7:
        sc[0], sc[1], sc[2] = 0, 1, R_MOD - 1
    pp = PlonkParams(ctx, ctx.srs_generate(fr_to_mont1(tau), n))
    got = pp.commit(z.Coefficients(fr_to_mont(sc))).affine()
    want = entry["commitment"]
    assert got == (None if want is None else (int(want[0], 16), int(want[1], 16)))
    # the same through msm_curve_addition on device-resident scalars
    buf = ctx.upload(fr_to_mont(sc))
    assert g1_from_mont(ctx.msm_dev(pp.srs, buf, 0, n)) == got


BUILD = {"range": lambda: circuits.range_circuit((1 << 64) - 1), "readme": circuits.readme_circuit,
         "logic_curve": circuits.logic_curve_circuit}


@pytest.mark.parametrize("entry", GOLD["proofs"], ids=lambda e: e["circuit"])
def test_proof_bytes_match_golden(ctx, entry):
    circ = SynthesizedCircuit.from_composer(BUILD[entry["circuit"]]())
    assert (circ.m, circ.n) == (entry["m"], entry["n"])
    rng = SplitMix64(entry["seed"])
    tau = rng.fr()
    pp = PlonkParams(ctx, ctx.srs_generate(fr_to_mont1(tau), entry["srs_len"]))
    prover = z.PlonkKey.compile_with_circuit(pp, entry["label"].encode(), circ)
    for nm, want in entry["verifier_key"].items():
        assert prover.verifier_key[nm] == (None if want is None else (int(want[0], 16), int(want[1], 16))), nm
    bl = [rng.fr() for _ in range(11)]
    assert ["%064x" % b for b in bl] == entry["blinders"]
    proof, pi = prover.create_proof(bl, circ)                  # native driver, witness gather on the device
    assert ["%064x" % p for p in pi] == entry["public_inputs"]
    assert proof.wire_bytes.hex() == entry["proof_bytes"]
    assert proof.to_bytes().hex() == entry["proof_bytes"]
    prover.native = False                                       # the same rounds driven from Python
    proof2, _ = prover.create_proof(bl, circ)
    assert proof2.to_bytes().hex() == entry["proof_bytes"]

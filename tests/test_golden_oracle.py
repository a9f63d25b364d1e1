"""The oracle still produces the committed golden vectors (tests/golden/vectors.json): guards the
checker against silent drift.  The quick subset runs here; `python tests/golden/make_golden.py`
regenerates everything."""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, json.load(open(os.path.join(HERE, "golden", "vectors.json")))


def test_oracle_reproduces_golden_ntt_and_msm():
    mod, gold = _load()
    got = mod.ntt_vectors(quick=True)
    want = [e for e in gold["ntt"] if e["k"] <= 12]
    assert got == want
    got = mod.msm_vectors(quick=True)
    want = [e for e in gold["msm"] if e["n"] <= 300]
    assert got == want


def test_oracle_reproduces_golden_proof():
    mod, gold = _load()
    got = mod.proof_vectors(quick=True)
    assert got == [e for e in gold["proofs"] if e["circuit"] == "range"]


def test_golden_constants_are_the_published_ones():
    """Public BLS12-381 facts the fixtures rest on, independent of any code in this repository."""
    _, gold = _load()
    # n = 1 MSM with scalar s over [tau^0]G = s * G; the range proof is 1040 bytes = 11 x 48 + 16 x 32
    assert all(len(p["proof_bytes"]) == 2 * 1040 for p in gold["proofs"])
    # every compressed commitment in a proof has the compression bit set (zcash encoding)
    for p in gold["proofs"]:
        raw = bytes.fromhex(p["proof_bytes"])
        assert all(raw[48 * i] & 0x80 for i in range(11))

"""Generates tests/golden/vectors.json with the CPU oracle (oracle/, Python big integers) on seeded
inputs.  Run here, in the build container:  python tests/golden/make_golden.py

The reference itself cannot be built or imported (Rust, absent sibling crates -- DESIGN.md section 2),
so these are the oracle's outputs frozen at a known-good state, not outputs of the reference.  They
serve two purposes: the GPU tests compare the CUDA path against them WITHOUT importing the oracle
(tests/test_gpu_golden.py), and a CPU test re-derives them so that the oracle cannot drift silently
(tests/test_golden_oracle.py).  Everything is hex of canonical integers / wire bytes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import curve, plonk as oplonk                       # noqa: E402
from oracle.fields import R_MOD                                  # noqa: E402
from oracle.ntt import Fft                                       # noqa: E402
from oracle.rng import SplitMix64                                # noqa: E402


def hx(v):
    return "%064x" % v


def pt(p):
    return None if p is None else ["%096x" % p[0], "%096x" % p[1]]


def digest(vals):
    """sha256 over the 32-byte little-endian encodings: what large vectors are stored as."""
    import hashlib
    h = hashlib.sha256()
    for x in vals:
        h.update(int(x).to_bytes(32, "little"))
    return h.hexdigest()


def ntt_vectors(quick=False):
    """Inputs are SplitMix64(seed).fr() draws (oracle/rng.py); small transforms are stored in full,
    large ones as digests."""
    out = []
    for k, len_in, seed in ((0, 1, 1), (1, 2, 2), (4, 16, 3), (6, 50, 4), (9, 512, 5), (12, 515, 6), (16, 65539, 7)):
        if quick and k > 12:
            continue
        rng = SplitMix64(seed)
        v = [rng.fr() for _ in range(len_in)]
        f = Fft(k)
        res = {"dft": f.dft(v), "idft": f.idft(v), "coset_dft": f.coset_dft(v), "coset_idft": f.coset_idft(v)} \
            if len_in <= (1 << k) else {"coset_dft_8n": Fft(k + 3).coset_dft(v)}
        e = {"k": k, "len_in": len_in, "seed": seed}
        for nm, r in res.items():
            if len(r) <= 64:
                e[nm] = [hx(x) for x in r]
            else:
                e[nm + "_sha256"] = digest(r)
        out.append(e)
    return out


def msm_vectors(quick=False):
    out = []
    # scalars: SplitMix64(seed): tau first, then n draws; for n >= 7 the first three are 0, 1, r - 1
    for n, seed in ((1, 11), (7, 12), (64, 13), (300, 14), (4096, 15)):
        if quick and n > 300:
            continue
        rng = SplitMix64(seed)
        tau = rng.fr()
        sc = [rng.fr() for _ in range(n)]
        if n >= 7:          # edge values the reference's commit sees: zero, one, r - 1
            sc[0], sc[1], sc[2] = 0, 1, R_MOD - 1
        dl, t = [], 1
        for _ in range(n):
            dl.append(t)
            t = t * tau % R_MOD
        out.append({"n": n, "seed": seed, "tau": hx(tau), "commitment": pt(curve.commit_known_dlog(dl, sc))})
    return out


def proof_vectors(quick=False):
    import circuits
    from host_mirror.composer import SynthesizedCircuit
    from oracle.merlin import Transcript
    out = []
    for name, build, label in (("range", lambda: circuits.range_circuit((1 << 64) - 1), b"demo"),
                               ("readme", circuits.readme_circuit, b"demo"),
                               ("logic_curve", circuits.logic_curve_circuit, b"plonk")):
        if quick and name != "range":
            continue
        circ = SynthesizedCircuit.from_composer(build())
        rng = SplitMix64(8349)
        tau = rng.fr()
        k = circ.n.bit_length() - 1
        srs_n = (1 << (max(k, 4) + 1)) + 7
        commit = oplonk.default_commit(tau=tau)
        opk, ovk = oplonk.compile_circuit(circ, commit, srs_n)
        tr = Transcript.base(label, oplonk.vk_transcript_list(ovk), circ.m)
        bl = [rng.fr() for _ in range(11)]
        proof, pi = oplonk.create_proof(opk, circ, commit, tr, bl)
        assert oplonk.verify(ovk, circ.n, proof, circ.pi_indexes, pi, tr, oplonk.trapdoor_kzg_check(tau))
        from dusk_plonk_b200.prover import Proof as HostProof
        hp = HostProof()
        for c in oplonk.Proof.COMM_NAMES:
            setattr(hp, c, getattr(proof, c))
        hp.evaluations = dict(proof.evaluations)
        out.append({"circuit": name, "label": label.decode(), "seed": 8349, "m": circ.m, "n": circ.n,
                    "srs_len": srs_n, "tau": hx(tau), "blinders": [hx(b) for b in bl],
                    "public_inputs": [hx(p) for p in pi],
                    "verifier_key": {nm: pt(ovk[nm]) for nm in list(oplonk.SELECTORS) +
                                     ["s_sigma_%d" % i for i in (1, 2, 3, 4)]},
                    "proof_bytes": hp.to_bytes().hex()})
    return out


def main():
    vec = {"generator": "tests/golden/make_golden.py (oracle/, Python big integers)",
           "ntt": ntt_vectors(), "msm": msm_vectors(), "proofs": proof_vectors()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
    with open(path, "w") as f:
        json.dump(vec, f, indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Generates tests/golden/synthetic_proofs.json: sha256 of the 1040-byte proof the ORACLE (C restatement +
the oracle's own Merlin and serialiser; no product code, no GPU) produces for the seeded synthetic
2^k-gate circuits that bench.py proves (seed 8349: tau, then the 11 blinders).  bench.py asserts the GPU
proof bytes against these digests inside every run, tests/test_gpu_prover.py does the same for -m gpu.

    python tests/golden/make_synthetic_digests.py 16 20     # minutes of CPU time at 2^20
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synthetic_proofs.json")


def main():
    import bench
    bench.host_threads()
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for logn in [int(a) for a in sys.argv[1:]] or [16]:
        t0 = time.time()
        cp, circ, bl = bench.cpu_prover(logn)
        proof, _ = cp.create_proof(bl, circ)
        raw = proof.to_bytes()
        res[str(logn)] = {"sha256": hashlib.sha256(raw).hexdigest(), "m": circ.m, "seed": 8349, "label": "plonk",
                          "a_comm_x": "%096x" % proof.a_comm[0], "r_poly_eval": "%064x" % proof.evaluations["r_poly_eval"],
                          "generator": "oracle/cprover.py (C restatement), oracle/merlin.py"}
        print(logn, res[str(logn)]["sha256"], "%.1f s" % (time.time() - t0), flush=True)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

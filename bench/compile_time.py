"""Times PlonkKey.compile (key preprocessing, src/key.rs:63-327: 15 iNTT(n) + 15 commits + 16 coset
NTT(8n) + SRS window table) on one GPU for the synthetic circuits.  Prints one JSON line per size."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dusk_plonk_b200 as z  # noqa: E402
from host_mirror.composer import synthetic_circuit  # noqa: E402
from dusk_plonk_b200.field import fr_to_mont1  # noqa: E402
from dusk_plonk_b200.plonk_params import PlonkParams  # noqa: E402

ctx = z.Context(0)
for logn in [int(a) for a in sys.argv[1:]] or [16]:
    circ = synthetic_circuit(logn)
    t0 = time.perf_counter()
    pp = PlonkParams.setup_synthetic(ctx, logn, fr_to_mont1(0x1234567))
    ctx.sync()
    t1 = time.perf_counter()
    prover = z.PlonkKey.compile(pp, circ)     # warm-up (domain tables, scratch)
    ctx.sync()
    t2 = time.perf_counter()
    l0 = ctx.launches
    import cProfile, pstats, io
    pr = cProfile.Profile()
    pr.enable()
    prover = z.PlonkKey.compile(pp, circ)
    ctx.sync()
    pr.disable()
    t3 = time.perf_counter()
    sio = io.StringIO()
    pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(10)
    sys.stderr.write(sio.getvalue())
    print(json.dumps({"log2_gates": logn, "srs_setup_ms": (t1 - t0) * 1e3, "compile_first_ms": (t2 - t1) * 1e3,
                      "compile_ms": (t3 - t2) * 1e3, "gpu_launches": ctx.launches - l0}), flush=True)

#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 zkplonk hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload prove|msm|ntt|compile] [--logn L]
    python bench.py --impl reference ...        # CPU restatement on the host cores (rank 0 only)

A "step" is one pass of the hot path over one batch of synthetic input: by default ONE
``create_proof`` of the synthetic 2^20-gate circuit (BASELINE config 5).  With N > 1 the driver
launches one rank per GPU through torchrun and the SAME job is sharded over the N GPUs (strong
scaling; SURVEY 8e: commits split by SRS ranges with a partial-sum gather, coset transforms and
quotient slices dealt out per GPU, four-step NTT with an all-to-all for --workload ntt); every rank
must end with byte-identical proofs, which is asserted inside the run together with the committed
digest of the same proof (tests/golden/synthetic_proofs.json).  ``--independent`` restores one
independent job per GPU (weak scaling, no data-path collective).  The 2^16-gate latency (BASELINE
config 2) rides along as the extra key ``prove16``.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# SURVEY 8(d) algorithmic work per element
MSM_IMAD_PER_POINT = 48000          # 16 signed 16-bit windows x 10 Fq mul x 300 mul-adds
NTT_BYTES_PER_ELEM = 64             # 32 B read + 32 B write, single pass lower bound
NTT_IMAD_PER_ELEM_PER_STAGE = 68    # 136 mul-adds per butterfly, N/2 butterflies per stage
# the dominant unit of work is one commit group's bucket accumulation, timed as one CUDA-event span ("a launch"):
ACC_GROUP = ("MSM accumulate group of one commit batch: msm_affine_first_kernel + msm_affine_round_kernel (+ their "
             "msm_affine_plan_kernel) for jobs >= 4 M pairs, then msm_accumulate_kernel (XYZZ tail / small jobs)")
PORT_NOTE = ("the port is a plain unsigned __int128 C restatement: about an order of magnitude per core below "
             "asm-backed CPU libraries; a reported baseline, not a target")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nme, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, logn):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this
    workload/size (profiles/traffic.json: bytes + the capture file it was read from), or None.  A capture
    older than the kernel's source file describes a different kernel and is refused."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, "no capture"
    ent = json.load(open(p)).get("%s:%d" % (workload, logn))
    if not ent:
        return None, "no capture for this workload/size"
    src = os.path.join(ROOT, ent.get("kernel_source", ""))
    if ent.get("kernel_source_sha256") and os.path.exists(src):
        import hashlib
        if hashlib.sha256(open(src, "rb").read()).hexdigest() != ent["kernel_source_sha256"]:
            return None, "capture %s predates the current %s: refused" % (ent.get("capture"), ent["kernel_source"])
    return ent["dram_bytes_per_launch"], ent.get("capture")


def imad_peak():
    """Integer multiply-add peak: measured by bench/imad_peak.cu if a result is committed
    under profiles/, else nominal 148 SM x 64 lanes x 1.965 GHz."""
    p = os.path.join(ROOT, "profiles", "imad_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["imad_peak_Tops"], "measured (profiles/imad_peak.json)"
    return 148 * 64 * 1.965e9 / 1e12, "nominal 148 SM x 64 lanes x 1.965 GHz"


# ---------------------------------------------------------------------------- reference arm
def host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is one process that should use the
    box's cores.  Must run before the oracle's OpenMP library is loaded."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


def cpu_prover(logn):
    """CPU restatement of compile + the witness for the same synthetic circuit.  Oracle side only:
    the circuit comes from the shared workload generator (host_mirror, pure numpy), the transcript is
    the oracle's own Merlin -- nothing of the product (and no libzkp_b200.so) is loaded."""
    from oracle import cport, cprover, curve
    from oracle.fields import R_MOD, fr_to_raw_limbs, g1_to_mont_limbs
    from oracle.merlin import Transcript
    from oracle.rng import SplitMix64
    from host_mirror.composer import synthetic_circuit
    circ = synthetic_circuit(logn)
    rng = SplitMix64(8349)
    tau = rng.fr()
    dl, t = [], 1
    for _ in range((1 << logn) + 7):
        dl.append(t)
        t = t * tau % R_MOD
    srs = cport.fixed_base_mul(g1_to_mont_limbs([curve.G1_GEN])[0], fr_to_raw_limbs(dl))
    cp = cprover.CProver(circ, srs, b"plonk", Transcript)
    bl = [rng.fr() for _ in range(11)]
    return cp, circ, bl


def run_reference(args, rank, world):
    """The reference's own CPU path cannot be built (Rust, absent crates): this arm times the
    oracle's threaded C restatement on a bounded sample of the same workload, on rank 0 alone."""
    if rank != 0:
        return
    cores_set = host_threads()
    from oracle import cport
    from oracle import curve
    from oracle.fields import R_MOD, fr_to_raw_limbs, g1_to_mont_limbs
    from oracle.rng import random_fr_raw_limbs
    cport.build()
    cores = cport.num_threads()
    assert cores == cores_set, (cores, cores_set)
    steps, warmup = args.steps, args.warmup
    if args.workload in ("prove", "compile"):
        t_c0 = time.perf_counter()
        cp, circ, bl = cpu_prover(args.logn)
        compile_s = time.perf_counter() - t_c0
        if args.workload == "compile":
            fn = None
            steps, warmup = 1, 0
            dt = compile_s
            unit, metric, n = "compiles/s", "compile_throughput", 1
            sample = ("SRS generation + PlonkKey::compile of the 2^%d-gate synthetic circuit, once; C restatement "
                      "(OpenMP) driven from Python" % args.logn)
        else:
            fn = lambda: cp.create_proof(bl, circ)
            # bounded: a 2^20-gate proof takes tens of seconds on the host cores
            steps = min(steps, 3 if args.logn <= 16 else 1)
            warmup = min(warmup, 1 if args.logn <= 16 else 0)
            unit, metric, n = "proofs/s", "create_proof_throughput", 1
            sample = ("one full create_proof of the 2^%d-gate synthetic circuit per step (%d step(s)), C restatement, "
                      "OpenMP (tuned: batch inversion, threaded loops)" % (args.logn, steps))
    elif args.workload == "msm":
        logs = min(args.logn, 16)
        n = 1 << logs
        g = g1_to_mont_limbs([curve.G1_GEN])[0]
        tau = 0x1234567890ABCDEF1234567890ABCDEF % (1 << 250)
        dl, t = [], 1
        for _ in range(n):
            dl.append(t); t = t * tau % R_MOD
        bases = cport.fixed_base_mul(g, fr_to_raw_limbs(dl))     # SRS-shaped bases tau^i G
        sc = random_fr_raw_limbs(8349, n)
        fn = lambda: cport.msm_g1(bases, sc)
        unit, metric = "Melem/s", "g1_msm_throughput"
        sample = "G1 MSM of 2^%d SRS-shaped points per step (workload size 2^%d), C restatement, OpenMP" % (logs, args.logn)
    else:
        logs = min(args.logn, 22)
        n = 1 << logs
        data = random_fr_raw_limbs(8349, n)
        fn = lambda: cport.ntt(data, logs)
        unit, metric = "Melem/s", "fr_ntt_throughput"
        sample = "Fr NTT of 2^%d elements per step (workload size 2^%d), C restatement, OpenMP" % (logs, args.logn)
    if fn is not None:
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t0) / steps
    val = n / dt if unit.endswith("s/s") and n == 1 else n / dt / 1e6
    line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak" if (args.independent and world > 1) else "strong", "vs_baseline": None,
            "dtype": "u64 limbs (modular integer)", "data": "synthetic", "config": workload_config(args, world),
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                             "note": PORT_NOTE},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "omp_threads": cores}
    if args.workload == "prove":
        line["prove_ms"] = dt * 1e3
        line["setup_s"] = compile_s
    print(json.dumps(line), flush=True)


def workload_config(args, world=1):
    shard = world > 1 and not args.independent
    if shard:
        par = {"prove": "ONE proof over %d GPUs, native driver: commits sharded by SRS ranges (partial sums gathered over "
                        "NCCL), the 8n quotient domain split by cosets (n-point transforms, point-wise quotient, one "
                        "slab exchange for the inverse transform)" % world,
               "compile": "ONE compile over %d GPUs, commits sharded by SRS ranges" % world,
               "msm": "ONE MSM over %d GPUs: SRS ranges + 96-byte partial-sum all-gather" % world,
               "ntt": "ONE transform over %d GPUs: four-step NTT (twiddle fused into the transpose), one NCCL "
                      "all-to-all issued by the library on its own stream" % world}[args.workload]
    else:
        par = "%d independent job(s), one per GPU, no data-path collective" % world
    if args.workload in ("prove", "compile"):
        what = "create_proof" if args.workload == "prove" else "PlonkKey::compile (15 iNTT + 15 commits + 16 coset NTT(8n))"
        return {"workload": "%s, synthetic 2^%d-gate add/mul circuit (n = 2^%d, quotient domain 8n = 2^%d), "
                            "11 commits + 11 NTT(n) + 8 NTT(8n) + element-wise rounds" % (what, args.logn, args.logn, args.logn + 3),
                "log2_gates": args.logn, "parallelism": par,
                "l2": "proving key + workspace (%d MiB) stream through each proof: larger than L2" %
                      ((16 + 9) * (1 << (args.logn + 3)) * 32 >> 20)}
    if args.workload == "msm":
        return {"workload": "G1 MSM (KZG commit) over 2^%d SRS powers, uniform scalars" % args.logn,
                "log2_n": args.logn, "parallelism": par, "l2": "inputs (bases+scalars) larger than L2"}
    return {"workload": "Fr NTT, 2^%d elements, forward, natural order" % args.logn, "log2_n": args.logn,
            "parallelism": par, "l2": "input larger than L2"}


def golden_digest(logn):
    p = os.path.join(ROOT, "tests", "golden", "synthetic_proofs.json")
    if os.path.exists(p):
        return json.load(open(p)).get(str(logn), {}).get("sha256")
    return None


# ------------------------------------------------------------------------------- GPU arm
class ProveJob:
    """compile + witness for the synthetic 2^logn-gate circuit on this rank's GPU; ``comm`` shards it."""

    def __init__(self, z, ctx, logn, comm=None):
        from host_mirror.composer import synthetic_circuit
        from host_mirror.synthetic import SplitMix64
        from dusk_plonk_b200.field import fr_to_mont1
        from dusk_plonk_b200.plonk_params import PlonkParams
        self.z, self.ctx, self.logn = z, ctx, logn
        circ = self.circ = synthetic_circuit(logn)
        rng = SplitMix64(8349)
        tau = rng.fr()
        t0 = time.perf_counter()
        if comm is not None:   # native multi-GPU driver: zkp_comm (NCCL inside libzkp_b200.so)
            from dusk_plonk_b200.plonk_params import ShardedNativeParams
            pp = ShardedNativeParams.setup_synthetic(ctx, comm, logn, fr_to_mont1(tau))
        else:
            pp = PlonkParams.setup_synthetic(ctx, logn, fr_to_mont1(tau))
        ctx.sync()
        self.srs_ms = (time.perf_counter() - t0) * 1e3
        self.pp = pp
        t0 = time.perf_counter()
        self.prover = z.PlonkKey.compile(pp, circ)
        ctx.sync()
        self.compile_ms = (time.perf_counter() - t0) * 1e3
        self.bl = [rng.fr() for _ in range(11)]
        self.sharded = comm is not None
        # witness values in pinned host memory; the wire gather runs on the device (every rank of a sharded
        # proof ships the same values to its own GPU)
        self.wa_host = z.WitnessValues.from_circuit(circ)
        self.wa_host.witness_mont = pinned_copy(z, self.wa_host.witness_mont)
        # (a rank of a sharded proof uploads 1 / G of the witness; the ranks all-gather the slices over NVLink)
        g = comm.world if comm is not None else 1
        self.h2d = (-(-self.wa_host.witness_mont.shape[0] // g) + len(self.wa_host.pi_values) + 19) * 32
        self.wa_dev = z.WitnessAssignment.from_circuit(circ, circ.n).to_device(ctx)
        self.d2h = 11 * 96 + 17 * 32
        self.proofs = []

    def step(self):
        self.proofs.append(self.prover.create_proof(self.bl, self.wa_dev)[0])

    def e2e_step(self):
        self.proofs.append(self.prover.create_proof(self.bl, self.wa_host)[0])

    def digest(self):
        import hashlib
        raws = [getattr(p, "wire_bytes", None) or p.to_bytes() for p in self.proofs]
        assert all(r == raws[0] for r in raws), "non-deterministic proofs"
        return hashlib.sha256(raws[0]).hexdigest()


def pinned_copy(z, a):
    p = z.pinned_empty(a.shape, a.dtype)
    p[...] = a
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="prove", choices=["prove", "msm", "ntt", "compile"])
    ap.add_argument("--logn", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prove16", action="store_true", help="skip the extra 2^16-gate latency key")
    ap.add_argument("--independent", action="store_true",
                    help="N > 1: one independent job per GPU (weak scaling) instead of ONE job sharded over the GPUs")
    ap.add_argument("--shard", action="store_true", help="(default for N > 1; kept for old command lines)")
    args = ap.parse_args()
    if args.logn is None:
        args.logn = {"prove": 20, "msm": 24, "ntt": 24, "compile": 20}[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import dusk_plonk_b200 as z
    from host_mirror.synthetic import random_fr_raw_limbs   # workload generator (numpy)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = z.Context(local_rank)
    n = 1 << args.logn
    hbm_peak, hbm_src = measured_peaks()
    imad_pk, imad_src = imad_peak()
    shard = world > 1 and not args.independent
    comm = ncomm = None
    if shard:
        import torch
        from dusk_plonk_b200.sharding import Communicator, FourStepNtt, ShardedPlonkParams
        comm = Communicator(torch.device("cuda", local_rank))
        if args.workload in ("prove", "compile", "ntt"):
            ncomm = z.NativeComm.from_torch_distributed(ctx)
    # seeds: independent jobs differ per rank, a sharded job is the same job on every rank
    jr = 0 if shard or world == 1 else rank

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    job = None
    extra = {}
    if args.workload == "compile":
        # a13: SRS window table + key preprocessing; a step is one whole compile (device-resident key)
        from host_mirror.composer import synthetic_circuit
        from host_mirror.synthetic import SplitMix64
        from dusk_plonk_b200.field import fr_to_mont1
        from dusk_plonk_b200.plonk_params import PlonkParams
        circ = synthetic_circuit(args.logn)
        taum = fr_to_mont1(SplitMix64(8349).fr())
        t0 = time.perf_counter()
        from dusk_plonk_b200.plonk_params import ShardedNativeParams
        pp = (ShardedNativeParams.setup_synthetic(ctx, ncomm, args.logn, taum) if shard
              else PlonkParams.setup_synthetic(ctx, args.logn, taum))
        ctx.sync()
        extra["srs_setup_ms"] = (time.perf_counter() - t0) * 1e3
        keep = []

        def step():
            keep.clear()
            keep.append(z.PlonkKey.compile(pp, circ))
        e2e_step = step
        # host -> device: 11 selector columns + sigma encodings; device -> host: 15 commitments
        h2d, d2h = 11 * circ.n * 32 + 4 * circ.n * 4, 15 * 96
        dominant, metric, n = "msm_accumulate", "compile_throughput", 1
    elif args.workload == "prove":
        job = ProveJob(z, ctx, args.logn, ncomm)
        step, e2e_step = job.step, job.e2e_step
        # whole-job bytes per step: every rank of a sharded proof ships its slice of the witness and reads the proof back
        h2d, d2h = job.h2d * (world if shard else 1), job.d2h * (world if shard else 1)
        extra["srs_setup_ms"], extra["compile_ms"] = job.srs_ms, job.compile_ms
        dominant, metric, n = "msm_accumulate", "create_proof_throughput", 1
    elif args.workload == "msm":
        tau = random_fr_raw_limbs(4242 + jr, 1)[0]
        host_scalars = pinned_copy(z, random_fr_raw_limbs(8349 + jr, n))
        dev_scalars = ctx.upload(host_scalars)
        if shard:
            lo, hi = z.sharding.shard_range(n, rank, world)
            sp = ShardedPlonkParams(ctx, comm, n, lo, hi, ctx.srs_generate(tau, hi - lo, first=lo))
            step = lambda: sp.commit(dev_scalars)

            def e2e_step():
                dev_scalars.upload(host_scalars[sp.lo:sp.hi], sp.lo)   # each rank ships only its range
                return sp.commit(dev_scalars)
            h2d, d2h = n * 32 // world, 96
        else:
            srs = ctx.srs_generate(tau, n)
            step = lambda: ctx.msm_dev(srs, dev_scalars, 0, n)
            e2e_step = lambda: ctx.msm(srs, host_scalars)
            h2d, d2h = n * 32, 96
        dominant, metric = "msm_accumulate", "g1_msm_throughput"
    else:
        host_data = pinned_copy(z, random_fr_raw_limbs(8349 + jr, n))
        if shard:
            fs = FourStepNtt(ctx, ncomm, args.logn)      # all-to-all by the library's own NCCL communicator
            fs.scatter_input(host_data)
            step = lambda: fs.run()
            host_out = z.pinned_empty((n, 4), np.uint64)

            def e2e_step():
                fs.scatter_input(host_data)
                fs.run()
                fs.gather_output(host_out)
            h2d, d2h = n * 32 // world, n * 32 // world
        else:
            src = ctx.upload(host_data)
            dst = ctx.alloc(n)
            step = lambda: ctx.ntt_dev(src, n, dst, args.logn, False, False)

            def e2e_step():   # in place on the pinned host vector, like Fft::dft by value
                ctx.check(ctx.lib.zkp_ntt(ctx.h, host_data.ctypes.data, n, args.logn, 0, 0))
            h2d, d2h = n * 32, n * 32
        dominant, metric = "ntt", "fr_ntt_throughput"

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.prof_enable(True)
    ctx.prof_reset()
    l0 = ctx.launches
    pts0 = ctx.msm_points
    comm0 = ncomm.stats() if ncomm is not None else (0, 0)
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop_ms()
    launches = ctx.launches - l0
    msm_pts = ctx.msm_points - pts0
    if ncomm is not None:
        c1 = ncomm.stats()
        extra["nccl"] = {"collectives_per_step": (c1[0] - comm0[0]) / args.steps,
                         "bytes_sent_per_rank_per_step": (c1[1] - comm0[1]) / args.steps,
                         "where": "inside libzkp_b200.so (zkp_comm): partial-commitment gathers, the coset exchange of "
                                  "the quotient's inverse transform, the gather of t(X)"}
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dom_ms, dom_cnt = ctx.prof_read(dominant)
    groups = ("msm_sort", "msm_accumulate", "msm_reduce", "msm_combine", "ntt", "quotient", "perm_z", "poly_eval",
              "poly_lincomb", "poly_div")
    prof_groups = {g: ctx.prof_read(g) for g in groups}
    ctx.prof_enable(False)
    ctx.prof_reset()

    # end to end through the host-buffer C-ABI call (H2D + kernels + D2H inside the timing)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    ctx.sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    if dist is not None:
        import torch
        t = torch.tensor([ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    # in-run parity: every rank holds byte-identical proofs, equal to the committed digest of this circuit
    if job is not None:
        dg = job.digest()
        digests = [dg]
        if dist is not None:
            digests = [None] * world
            dist.all_gather_object(digests, dg)
        if shard:
            assert all(d == digests[0] for d in digests), "ranks disagree on the proof bytes: %r" % (digests,)
        gold = golden_digest(args.logn) if (shard or world == 1) else None
        assert gold is None or dg == gold, "proof bytes differ from tests/golden/synthetic_proofs.json"
        extra["proof_sha256"] = dg
        extra["proof_check"] = ("%d rank(s) byte-identical" % len(digests) if shard or world == 1 else "independent proofs") + \
            ("; equals the committed digest (GPU == CPU oracle, tests/golden/synthetic_proofs.json)" if gold else
             "; no committed digest for this size")
        # BASELINE config 2 rides along: latency of a 2^16-gate proof on one GPU (rank 0)
        if not args.no_prove16 and args.logn != 16 and rank == 0:
            ctx16 = z.Context(local_rank)
            j16 = ProveJob(z, ctx16, 16)
            for _ in range(3):
                j16.step()
            ctx16.sync()
            ctx16.timer_start()
            for _ in range(10):
                j16.step()
            ms16 = ctx16.timer_stop_ms() / 10
            j16.e2e_step()
            ctx16.sync()
            t0 = time.perf_counter()
            for _ in range(10):
                j16.e2e_step()
            ctx16.sync()
            e16 = (time.perf_counter() - t0) * 1e2
            # Prover: Clone (src/prover.rs:28): several provers over one key on one GPU, one host thread each
            def concurrent(nprov, per):
                import threading
                provs = [j16.prover] + [j16.prover.clone() for _ in range(nprov - 1)]
                outs = [[] for _ in provs]

                def work(p, o, cnt):
                    for _ in range(cnt):
                        o.append(p.create_proof(j16.bl, j16.wa_host)[0])
                for p, o in zip(provs, outs):       # warm-up: scratch, domain tables, wiring of every clone
                    work(p, o, 2)
                for p in provs:
                    p.ctx.sync()
                th = [threading.Thread(target=work, args=(p, o, per)) for p, o in zip(provs, outs)]
                t0 = time.perf_counter()
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                for p in provs:
                    p.ctx.sync()
                dt = time.perf_counter() - t0
                for o in outs:
                    j16.proofs.extend(o)
                for p in provs[1:]:
                    p.close()
                return nprov * per / dt
            conc = {str(kc): concurrent(kc, 20) for kc in (2, 3)}
            d16, g16 = j16.digest(), golden_digest(16)
            assert g16 is None or d16 == g16, "2^16 proof bytes differ from the committed digest"
            extra["prove16"] = {"prove_ms": ms16, "e2e_prove_ms": e16, "proofs_per_s": 1e3 / ms16, "n_gpus": 1,
                                "concurrent_provers_e2e_proofs_per_s": conc,
                                "concurrent_note": "k cloned provers (Prover: Clone) on one GPU, one host thread each, pinned "
                                                   "host witness in, proof bytes out; batch-1 latency is prove_ms",
                                "proof_sha256": d16, "equals_golden": g16 is not None}
            j16.prover.close()
            ctx16.close()
        if dist is not None:
            dist.barrier()

    if rank == 0:
        ms_step = ms / args.steps
        jobs = 1 if (shard or world == 1) else world      # sharded: ONE job over all GPUs (strong scaling)
        value = jobs * n / (ms_step * 1e-3) / 1e6
        e2e_val = jobs * n / (e2e_ms / args.steps * 1e-3) / 1e6
        dom_avg_ms = dom_ms / max(dom_cnt, 1)
        unit = "Melem/s"
        if args.workload in ("prove", "compile"):
            unit = "proofs/s" if args.workload == "prove" else "compiles/s"
            value, e2e_val = value * 1e6, e2e_val * 1e6
            # dominant kernel: msm_accumulate over the commits of a step (this rank's share when sharded)
            achieved = MSM_IMAD_PER_POINT * (msm_pts / max(dom_cnt, 1)) / (dom_avg_ms * 1e-3) / 1e12
            shares = {}
            for g in groups:
                gm, gc = prof_groups.get(g, (0.0, 0))
                shares[g] = {"ms_per_step": gm / args.steps, "launch_groups_per_step": gc / args.steps}
            roofline = {"bound": "imad", "kernel": ACC_GROUP, "achieved": achieved, "peak": imad_pk,
                        "unit": "T IMAD/s", "frac": achieved / imad_pk, "traffic": None, "peak_source": imad_src,
                        "kernel_ms": dom_avg_ms, "launches_per_step": dom_cnt / args.steps,
                        "kernel_share_of_step": dom_ms / args.steps / ms_step,
                        "algorithmic": "48000 mul-adds per committed point (SURVEY 8d); %d points per step on this GPU" %
                                       (msm_pts // args.steps)}
            extra.update({"prove_ms" if args.workload == "prove" else "compile_ms": ms_step,
                          "e2e_ms": e2e_ms / args.steps, "kernel_groups": shares})
        elif args.workload == "msm":
            # the roofline is per GPU: a sharded job's kernel on this rank sums n / world points
            n_local = n // world if shard else n
            achieved = MSM_IMAD_PER_POINT * n_local / (dom_avg_ms * 1e-3) / 1e12
            roofline = {"bound": "imad", "kernel": ACC_GROUP, "achieved": achieved, "peak": imad_pk,
                        "unit": "T IMAD/s", "frac": achieved / imad_pk, "traffic": None, "peak_source": imad_src,
                        "kernel_ms": dom_avg_ms, "kernel_share_of_step": dom_avg_ms / ms_step,
                        "algorithmic": "48000 mul-adds per point (SURVEY 8d)"}
        else:
            n_local = n // world if shard else n   # per-GPU roofline: this rank transforms n / world elements
            achieved = NTT_BYTES_PER_ELEM * n_local / (dom_avg_ms * 1e-3) / 1e9
            imad = NTT_IMAD_PER_ELEM_PER_STAGE * n_local * args.logn / (dom_avg_ms * 1e-3) / 1e12
            roofline = {"bound": "hbm", "kernel": "ntt_pass_kernel (all passes)", "achieved": achieved,
                        "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                        "peak_source": hbm_src, "kernel_ms": dom_avg_ms, "kernel_share_of_step": dom_avg_ms / ms_step,
                        "imad_achieved_T": imad, "imad_frac": imad / imad_pk, "imad_peak_source": imad_src,
                        "algorithmic": "64 B per element; 68*N*log2(N) mul-adds (SURVEY 8d)"}
        roofline["traffic"], roofline["traffic_source"] = measured_traffic(args.workload, args.logn)
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                # one job whatever N is (strong scaling) unless --independent gives every GPU its own job
                "scaling": "weak" if (args.independent and world > 1) else "strong",
                "vs_baseline": None, "dtype": "u32 limbs (modular integer)", "data": "synthetic",
                "config": workload_config(args, world),
                "roofline": roofline,
                "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "clocks": clocks}
        line.update(extra)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if job is not None:
        job.prover.close()
    if ncomm is not None:
        ncomm.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(args):
    """Oracle C restatement on the host cores, bounded sample (rank 0 only)."""
    cores_set = host_threads()
    from oracle import cport
    from oracle.rng import random_fr_raw_limbs
    cport.build()
    cores = cport.num_threads()
    if args.workload in ("prove", "compile"):
        # a 2^20-gate proof costs the host cores the better part of a minute (plus minutes of CPU key
        # set-up): the sample is one whole proof of the 2^16-gate circuit of the same family, scaled by the
        # gate ratio (MSM ~ n / log n and NTT ~ n log n straddle linear); `bench.py --impl reference` times
        # the full-size proof itself
        logs = min(args.logn, 16)
        scale = float(1 << (args.logn - logs))
        t_c0 = time.perf_counter()
        cp, circ, bl = cpu_prover(logs)
        t_c = time.perf_counter() - t_c0
        if args.workload == "compile":
            dt = t_c
            what = "SRS generation + compile"
        else:
            t0 = time.perf_counter()
            cp.create_proof(bl, circ)
            dt = time.perf_counter() - t0
            what = "create_proof"
        unit = "proofs/s" if args.workload == "prove" else "compiles/s"
        return {"value": 1.0 / (dt * scale), "unit": unit, "cores": cores, "kind": "port", "note": PORT_NOTE,
                "sample_ms": dt * 1e3, "sample_log2_gates": logs, "scaled_by": scale,
                "sample": "one full %s of the 2^%d-gate synthetic circuit%s; C restatement, OpenMP, tuned (batch "
                          "inversion, threaded loops)" % (what, logs, "" if scale == 1.0 else
                                                          ", time scaled x%d to 2^%d gates" % (scale, args.logn))}
    if args.workload == "msm":
        from oracle import curve
        from oracle.fields import R_MOD, fr_to_raw_limbs, g1_to_mont_limbs
        logs = min(args.logn, 16)
        n = 1 << logs
        dl, t = [], 1
        for _ in range(n):
            dl.append(t); t = t * 0x9E3779B97F4A7C15F39CC0605CEDC835 % R_MOD
        bases = cport.fixed_base_mul(g1_to_mont_limbs([curve.G1_GEN])[0], fr_to_raw_limbs(dl))
        sc = random_fr_raw_limbs(8349, n)
        cport.msm_g1(bases, sc)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            cport.msm_g1(bases, sc)
        dt = (time.perf_counter() - t0) / reps
        sample = "G1 MSM 2^%d points x %d reps (of the 2^%d workload)" % (logs, reps, args.logn)
    else:
        logs = min(args.logn, 22)
        n = 1 << logs
        data = random_fr_raw_limbs(8349, n)
        cport.ntt(data, logs)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            cport.ntt(data, logs)
        dt = (time.perf_counter() - t0) / reps
        sample = "Fr NTT 2^%d elements x %d reps (of the 2^%d workload)" % (logs, reps, args.logn)
    return {"value": n / dt / 1e6, "unit": "Melem/s", "cores": cores, "kind": "port", "sample": sample, "note": PORT_NOTE}


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 zkplonk hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload msm|ntt] [--logn L]
    python bench.py --impl reference ...        # CPU restatement on the host cores

A "step" is one pass of the hot path over one batch of synthetic input.  With N > 1 the
driver launches one rank per GPU through torchrun; the path shards by independent
commitments, so every rank processes its own batch (weak scaling) and the only
collective is the timing barrier / max-reduce.
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
    """DRAM bytes of the dominant kernel from the committed ncu --set full capture, if one exists for
    this workload/size (profiles/traffic.json); else None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        ent = json.load(open(p)).get("%s:%d" % (workload, logn))
        if ent:
            return ent["dram_bytes_per_launch"]
    return None


def imad_peak():
    """Integer multiply-add peak: measured by bench/imad_peak.cu if a result is committed
    under profiles/, else nominal 148 SM x 64 lanes x 1.965 GHz."""
    p = os.path.join(ROOT, "profiles", "imad_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["imad_peak_Tops"], "measured (profiles/imad_peak.json)"
    return 148 * 64 * 1.965e9 / 1e12, "nominal 148 SM x 64 lanes x 1.965 GHz"


# ---------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU path cannot be built (Rust, absent crates): this arm times the
    oracle's threaded C restatement on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import cport
    from oracle.fields import g1_to_mont_limbs
    from oracle import curve
    from oracle.rng import random_fr_raw_limbs
    cport.build()
    cores = cport.num_threads()
    if args.workload == "prove":
        cp, circ, bl = cpu_prover(args.logn)
        fn = lambda: cp.create_proof(bl, circ)
        n = 1
        unit, metric = "proofs/s", "create_proof_throughput"
        sample = ("one full create_proof of the 2^%d-gate synthetic circuit per step, C restatement, OpenMP "
                  "(tuned: batch inversion, threaded loops)" % args.logn)
        for _ in range(min(args.warmup, 1)):
            fn()
        steps = min(args.steps, 3)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t0) / steps
        val = 1.0 / dt
        line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": world,
                "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (modular integer)", "data": "synthetic",
                "config": workload_config(args), "prove_ms": dt * 1e3,
                "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    if args.workload == "msm":
        logs = min(args.logn, 16)
        n = 1 << logs
        g = g1_to_mont_limbs([curve.G1_GEN])[0]
        # SRS-shaped bases tau^i G are produced by the oracle's fixed-base routine
        from oracle.fields import fr_to_raw_limbs
        tau = 0x1234567890ABCDEF1234567890ABCDEF % (1 << 250)
        dl, t = [], 1
        from oracle.fields import R_MOD
        for _ in range(n):
            dl.append(t); t = t * tau % R_MOD
        bases = cport.fixed_base_mul(g, fr_to_raw_limbs(dl))
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
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt / 1e6
    line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (modular integer)", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_prover(logn):
    """CPU restatement of compile + the witness for the same synthetic circuit (oracle side)."""
    from oracle import cport, cprover, curve
    from oracle.fields import R_MOD, fr_to_raw_limbs, g1_to_mont_limbs
    from oracle.rng import SplitMix64
    from dusk_plonk_b200.composer import synthetic_circuit
    from dusk_plonk_b200.transcript import Transcript
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


def workload_config(args):
    if args.workload == "prove":
        return {"workload": "create_proof, synthetic 2^%d-gate add/mul circuit (n = 2^%d, quotient domain 8n = 2^%d), "
                            "11 commits + 11 NTT(n) + 8 NTT(8n) + element-wise rounds" % (args.logn, args.logn, args.logn + 3),
                "log2_gates": args.logn,
                "l2": "proving key + workspace (%d MiB) stream through each proof: larger than L2" %
                      ((16 + 9) * (1 << (args.logn + 3)) * 32 >> 20)}
    if args.workload == "msm":
        return {"workload": "G1 MSM (KZG commit) over 2^%d SRS powers, uniform scalars" % args.logn,
                "log2_n": args.logn, "l2": "inputs (bases+scalars) larger than L2"}
    return {"workload": "Fr NTT, 2^%d elements, forward, natural order" % args.logn, "log2_n": args.logn,
            "l2": "input larger than L2"}


# ------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="prove", choices=["prove", "msm", "ntt"])
    ap.add_argument("--logn", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", action="store_true",
                    help="N > 1: shard ONE job over the GPUs (SRS ranges + partial-sum gather for commits, four-step "
                         "NTT with all-to-all) and report strong scaling; default is one independent job per GPU")
    args = ap.parse_args()
    if args.logn is None:
        args.logn = {"prove": 16, "msm": 22, "ntt": 24}[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import dusk_plonk_b200 as z
    from dusk_plonk_b200.synthetic import random_fr_raw_limbs   # the product's own input generator

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
    shard = args.shard and world > 1
    comm = None
    if shard:
        import torch
        from dusk_plonk_b200.sharding import Communicator, FourStepNtt, ShardedPlonkParams
        comm = Communicator(torch.device("cuda", local_rank))
    # seeds: independent jobs differ per rank, a sharded job is the same job on every rank
    jr = 0 if shard else rank

    def pinned_copy(a):
        p = z.pinned_empty(a.shape, a.dtype)
        p[...] = a
        return p

    if args.workload == "prove":
        from dusk_plonk_b200.composer import synthetic_circuit
        from dusk_plonk_b200.field import fr_to_mont1
        from dusk_plonk_b200.plonk_params import PlonkParams
        from dusk_plonk_b200.synthetic import SplitMix64
        circ = synthetic_circuit(args.logn)
        rng = SplitMix64(8349)
        tau = rng.fr()
        if shard:
            pp = ShardedPlonkParams.setup_synthetic(ctx, comm, args.logn, fr_to_mont1(tau))
        else:
            pp = PlonkParams.setup_synthetic(ctx, args.logn, fr_to_mont1(tau))
        prover = z.PlonkKey.compile(pp, circ)
        bl = [rng.fr() for _ in range(11)]
        if shard:    # the sharded proof is driven round by round from Python: host-gathered wire columns
            wa_host = z.WitnessAssignment.from_circuit(circ, circ.n)
            wa_host.wires_mont = pinned_copy(wa_host.wires_mont)
            wa_host.dense_pi_mont = pinned_copy(wa_host.dense_pi_mont)
            h2d_prove = 5 * circ.n * 32 + 19 * 32
        else:        # witness values in pinned host memory; the wire gather runs on the device
            wa_host = z.WitnessValues.from_circuit(circ)
            wa_host.witness_mont = pinned_copy(wa_host.witness_mont)
            h2d_prove = (wa_host.witness_mont.shape[0] + len(wa_host.pi_values) + 19) * 32
        wa_dev = z.WitnessAssignment.from_circuit(circ, circ.n).to_device(ctx)
        proofs = []
        step = lambda: proofs.append(prover.create_proof(bl, wa_dev)[0])
        e2e_step = lambda: proofs.append(prover.create_proof(bl, wa_host)[0])
        h2d, d2h = h2d_prove, 11 * 96 + 17 * 32
        dominant = "msm_accumulate"
        metric = "create_proof_throughput"
        n = 1
    elif args.workload == "msm":
        tau = random_fr_raw_limbs(4242 + jr, 1)[0]
        host_scalars = pinned_copy(random_fr_raw_limbs(8349 + jr, n))
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
        dominant = "msm_accumulate"
        metric = "g1_msm_throughput"
    else:
        host_data = pinned_copy(random_fr_raw_limbs(8349 + jr, n))
        if shard:
            fs = FourStepNtt(ctx, comm, args.logn)
            fs.scatter_input(host_data)
            step = lambda: fs.run()
            host_out = np.zeros((n, 4), dtype=np.uint64)

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
        dominant = "ntt"
        metric = "fr_ntt_throughput"

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

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
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop_ms()
    launches = ctx.launches - l0
    msm_pts = ctx.msm_points - pts0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dom_ms, dom_cnt = ctx.prof_read(dominant)
    prof_groups = {g: ctx.prof_read(g) for g in ("msm_sort", "msm_accumulate", "msm_reduce", "msm_combine", "ntt",
                                                 "quotient", "perm_z", "poly_eval", "poly_lincomb", "poly_div")}
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

    if rank == 0:
        ms_step = ms / args.steps
        jobs = 1 if shard else world      # sharded: ONE job over all GPUs (strong scaling)
        value = jobs * n / (ms_step * 1e-3) / 1e6
        e2e_val = jobs * n / (e2e_ms / args.steps * 1e-3) / 1e6
        dom_avg_ms = dom_ms / max(dom_cnt, 1)
        unit = "Melem/s"
        extra = {}
        if args.workload == "prove":
            unit = "proofs/s"
            value, e2e_val = value * 1e6, e2e_val * 1e6
            assert all(p == proofs[0] for p in proofs), "non-deterministic proofs"
            # dominant kernel: msm_accumulate over the proof's 11 commits
            achieved = MSM_IMAD_PER_POINT * (msm_pts / max(dom_cnt, 1)) / (dom_avg_ms * 1e-3) / 1e12
            shares = {}
            for g in ("msm_sort", "msm_accumulate", "msm_reduce", "msm_combine", "ntt", "quotient", "perm_z",
                      "poly_eval", "poly_lincomb", "poly_div"):
                gm, gc = prof_groups.get(g, (0.0, 0))
                shares[g] = {"ms_per_proof": gm / args.steps, "launch_groups_per_proof": gc / args.steps}
            roofline = {"bound": "imad", "kernel": "msm_accumulate_kernel", "achieved": achieved, "peak": imad_pk,
                        "unit": "T IMAD/s", "frac": achieved / imad_pk, "traffic": None, "peak_source": imad_src,
                        "kernel_ms": dom_avg_ms, "launches_per_step": dom_cnt / args.steps,
                        "kernel_share_of_step": dom_ms / args.steps / ms_step,
                        "algorithmic": "48000 mul-adds per committed point (SURVEY 8d); %d points per proof" %
                                       (msm_pts // args.steps)}
            extra = {"prove_ms": ms_step, "e2e_prove_ms": e2e_ms / args.steps, "kernel_groups": shares}
        elif args.workload == "msm":
            # the roofline is per GPU: a sharded job's kernel on this rank sums n / world points
            n_local = n // world if shard else n
            achieved = MSM_IMAD_PER_POINT * n_local / (dom_avg_ms * 1e-3) / 1e12
            roofline = {"bound": "imad", "kernel": "msm_accumulate_kernel", "achieved": achieved, "peak": imad_pk,
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
        roofline["traffic"] = measured_traffic(args.workload, args.logn)
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if shard else "weak",
                "vs_baseline": None, "dtype": "u32 limbs (modular integer)", "data": "synthetic",
                "config": dict(workload_config(args), parallelism=(
                    "one job sharded over %d GPUs (SRS ranges + 96-byte partial-sum all-gather; four-step NTT + "
                    "all-to-all)" % world if shard else "%d independent job(s), one per GPU, no data-path collective" % world)),
                "roofline": roofline,
                "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "clocks": clocks}
        line.update(extra)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(args):
    """Oracle C restatement on the host cores, bounded sample (rank 0 only)."""
    from oracle import cport
    from oracle.rng import random_fr_raw_limbs
    cport.build()
    cores = cport.num_threads()
    if args.workload == "prove":
        cp, circ, bl = cpu_prover(args.logn)
        t0 = time.perf_counter()
        cp.create_proof(bl, circ)
        dt = time.perf_counter() - t0
        return {"value": 1.0 / dt, "unit": "proofs/s", "cores": cores, "kind": "port", "prove_ms": dt * 1e3,
                "sample": "one full create_proof of the same 2^%d-gate circuit (whole workload, not a sample); "
                          "C restatement, OpenMP, tuned (batch inversion, threaded loops)" % args.logn}
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
    return {"value": n / dt / 1e6, "unit": "Melem/s", "cores": cores, "kind": "port", "sample": sample}


if __name__ == "__main__":
    main()

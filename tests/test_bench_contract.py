"""CPU: pieces of bench.py's contract that need no GPU -- both arms describe the same workload, the committed ncu
capture behind ``roofline.traffic`` matches the current kernel sources (and is refused when it does not), the
oracle digests of the timed proofs are present, the reference arm loads no product library."""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(workload, logn, independent=False):
    return argparse.Namespace(workload=workload, logn=logn, independent=independent)


def test_both_arms_use_one_workload_description():
    for w, l in (("prove", 20), ("msm", 24), ("ntt", 26), ("compile", 16)):
        for world in (1, 2, 8):
            a, b = bench.workload_config(_args(w, l), world), bench.workload_config(_args(w, l), world)
            assert a == b and "workload" in a and "parallelism" in a
    assert "ONE proof over 8 GPUs" in bench.workload_config(_args("prove", 20), 8)["parallelism"]
    assert "independent" in bench.workload_config(_args("prove", 20, True), 8)["parallelism"]


def test_traffic_capture_matches_the_current_kernel_sources(tmp_path, monkeypatch):
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for key, ent in traffic.items():
        src = os.path.join(ROOT, ent["kernel_source"])
        assert hashlib.sha256(open(src, "rb").read()).hexdigest() == ent["kernel_source_sha256"], \
            "%s: re-capture with bench/r02_capture.sh + bench/make_traffic.py after changing %s" % (key, ent["kernel_source"])
        w, l = key.split(":")
        val, where = bench.measured_traffic(w, int(l))
        assert val == ent["dram_bytes_per_launch"] and where == ent["capture"]
    # a capture of other sources is refused, not reported
    fake = {k: dict(v, kernel_source_sha256="0" * 64) for k, v in traffic.items()}
    (tmp_path / "profiles").mkdir()
    (tmp_path / "profiles" / "traffic.json").write_text(json.dumps(fake))
    for k, v in traffic.items():
        dst = tmp_path / v["kernel_source"]
        dst.parent.mkdir(parents=True, exist_ok=True)
        dst.write_bytes(open(os.path.join(ROOT, v["kernel_source"]), "rb").read())
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    val, why = bench.measured_traffic("prove", 20)
    assert val is None and "refused" in why


def test_digests_of_the_timed_proofs_are_committed():
    for logn in (12, 16, 20):
        d = bench.golden_digest(logn)
        assert d and len(d) == 64


def test_reference_arm_maps_no_product_library():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'ntt', '--logn', '10', "
            "'--steps', '1', '--warmup', '0']\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "libs = sorted({l.split()[-1] for l in open('/proc/self/maps') if 'libzkp' in l})\n"
            "print('LIBS', libs)\n" % os.path.join(ROOT, "bench.py"))
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="4")     # as torchrun would set them
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["impl"] == "reference" and rec["n_gpus"] == 4
    assert rec["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) > 1 or os.cpu_count() == 1
    libs = [l for l in out.stdout.splitlines() if l.startswith("LIBS")][-1]
    assert "libzkp_oracle" in libs and "libzkp_b200" not in libs
    # the other ranks of a torchrun launch exit at once without output
    env["RANK"] = "3"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ntt",
                          "--logn", "10"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""

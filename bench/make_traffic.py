"""profiles/traffic.json from the committed ncu --set full summaries (profiles/r02/ncu_*.json): DRAM bytes per launch
of the dominant kernel GROUP of each bench workload, with the capture file and the sha256 of the kernel source the
capture was taken from -- bench.py refuses an entry whose source has changed since (a stale capture describes a
different kernel).  Usage: python bench/make_traffic.py"""
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(path):
    return hashlib.sha256(open(os.path.join(ROOT, path), "rb").read()).hexdigest()


def num(s):
    v, u = s.split()[0].replace(",", ""), s.split()[1] if len(s.split()) > 1 else ""
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


def group_bytes(path, first, count):
    rows = json.load(open(os.path.join(ROOT, path)))[first:first + count]
    return sum(num(r["dram__bytes_read.sum"]) + num(r["dram__bytes_write.sum"]) for r in rows), [r["Kernel Name"].split("(")[0] for r in rows]


out = {}
msm_src, ntt_src = "dusk-plonk_b200/csrc/msm.cu", "dusk-plonk_b200/csrc/ntt.cu"
b, ks = group_bytes("profiles/r02/ncu_acc_prove20.json", 0, 8)
out["prove:20"] = {"dram_bytes_per_launch": b, "capture": "profiles/r02/ncu_acc_prove20.json (launches 0-7: one accumulate "
                   "group = the four wire commitments: plan / first round / 3 later rounds / XYZZ tail)",
                   "kernels": ks, "kernel_source": msm_src, "kernel_source_sha256": sha(msm_src)}
b, ks = group_bytes("profiles/r02/ncu_acc_msm24.json", 0, 10)
out["msm:24"] = {"dram_bytes_per_launch": b, "capture": "profiles/r02/ncu_acc_msm24.json (launches 0-9: one MSM's accumulate "
                 "group: first round, 4 later rounds with their plans, XYZZ tail)",
                 "kernels": ks, "kernel_source": msm_src, "kernel_source_sha256": sha(msm_src)}
b, ks = group_bytes("profiles/r02/ncu_ntt24.json", 0, 3)
out["ntt:24"] = {"dram_bytes_per_launch": b, "capture": "profiles/r02/ncu_ntt24.json (the three passes of one transform)",
                 "kernels": ks, "kernel_source": ntt_src, "kernel_source_sha256": sha(ntt_src)}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
for k, v in out.items():
    print(k, "%.3f GB" % (v["dram_bytes_per_launch"] / 1e9), v["kernels"])

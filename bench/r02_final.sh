#!/bin/bash
# round 2: what the driver runs at round end, on one GPU: smoke(), the -m gpu suite, both bench arms
mkdir -p gpurun_out/r02
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02/final_smoke.log
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02/final_pytest_gpu.log; tail -2 gpurun_out/r02/final_pytest_gpu.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02/final_bench_n1.json 2> gpurun_out/r02/final_bench_n1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02/final_bench_n1.json'))
print('prove20 %.2f ms e2e %.2f | prove16 %.2f ms conc %s | frac %.3f traffic %s | cpu %s'%(d['prove_ms'], d['e2e_ms'], d['prove16']['prove_ms'], d['prove16']['concurrent_provers_e2e_proofs_per_s'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value']))"
if [ "$1" = "ref" ]; then
  python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02/final_bench_ref.json 2> gpurun_out/r02/final_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02/final_bench_ref.json
fi

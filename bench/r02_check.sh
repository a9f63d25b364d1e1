#!/bin/bash
# round 2: full GPU suite + default bench + sanitizer passes on the README-circuit proof (memcheck, then racecheck
# in a second call: one tool per gpurun call)
mkdir -p gpurun_out/r02
TOOL=${1:-none}
if [ "$TOOL" = "none" ]; then
  (timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02/pytest_gpu_b.log; tail -3 gpurun_out/r02/pytest_gpu_b.log
  python bench.py --steps 5 --warmup 3 > gpurun_out/r02/bench_prove20_b.json 2> gpurun_out/r02/bench_prove20_b.err
  python -c "
import json; d=json.load(open('gpurun_out/r02/bench_prove20_b.json'))
print('prove20 %.2f ms e2e %.2f ms | prove16 %.2f ms | frac %.3f | setup srs %.0f ms compile %.0f ms'%(d['prove_ms'], d['e2e_ms'], d['prove16']['prove_ms'], d['roofline']['frac'], d['srs_setup_ms'], d['compile_ms']))
print({k:round(v['ms_per_step'],2) for k,v in d['kernel_groups'].items()})"
else
  # the smallest case that runs every kernel of a proof: smoke() = README-size NTT, MSM, the range-circuit proof
  timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/sanitizer_$TOOL.log 2>&1
  echo "sanitizer $TOOL rc=$?"; tail -5 gpurun_out/r02/sanitizer_$TOOL.log
fi

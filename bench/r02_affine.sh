#!/bin/bash
# round 2: batched-affine accumulation -- correctness with the rounds forced on for every size, then A/B timings
mkdir -p gpurun_out/r02
(ZKP_MSM_AFFINE_MIN=1 timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden.py -x -q 2>&1 | tail -15) > gpurun_out/r02/pytest_affine_forced.log
(ZKP_MSM_AFFINE_MIN=1 ZKP_MSM_AFFINE_ROUNDS=3 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q 2>&1 | tail -5) >> gpurun_out/r02/pytest_affine_forced.log
tail -4 gpurun_out/r02/pytest_affine_forced.log
for R in 0 auto; do
  if [ $R = auto ]; then unset ZKP_MSM_AFFINE_ROUNDS; else export ZKP_MSM_AFFINE_ROUNDS=$R; fi
  for L in 18 20 22 24; do
    python bench.py --workload msm --logn $L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02/aff_msm${L}_R$R.json 2> gpurun_out/r02/aff_msm${L}_R$R.err
    python - <<PY
import json
d=json.load(open("gpurun_out/r02/aff_msm${L}_R$R.json"))
print("msm 2^$L R=$R: %.3f ms  frac %.3f  kernel %.3f ms"%(d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]))
PY
  done
  for L in 16 20; do
    python bench.py --workload prove --logn $L --steps 5 --warmup 3 --no-cpu-baseline --no-prove16 > gpurun_out/r02/aff_prove${L}_R$R.json 2> gpurun_out/r02/aff_prove${L}_R$R.err
    python - <<PY
import json
d=json.load(open("gpurun_out/r02/aff_prove${L}_R$R.json"))
print("prove 2^$L R=$R: %.3f ms  acc %.3f  " % (d["ms_per_step"], d["kernel_groups"]["msm_accumulate"]["ms_per_step"]), d["proof_check"][:40])
PY
  done
done

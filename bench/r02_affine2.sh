#!/bin/bash
# round 2: quick loop for the batched-affine kernel -- forced-rounds correctness, three timings, one ncu capture
mkdir -p gpurun_out/r02
(ZKP_MSM_AFFINE_MIN=1 timeout 900 python -m pytest tests/test_gpu_msm.py -x -q 2>&1 | tail -4) > gpurun_out/r02/pytest_affine_forced.log
tail -2 gpurun_out/r02/pytest_affine_forced.log
for W in "msm 22" "msm 24" "prove 20" "prove 16"; do
  set -- $W
  python bench.py --workload $1 --logn $2 --steps 5 --warmup 3 --no-cpu-baseline --no-prove16 > gpurun_out/r02/aff2_$1$2.json 2> gpurun_out/r02/aff2_$1$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02/aff2_$1$2.json"))
print("$1 2^$2: %.3f ms  frac %.3f  kernel %.3f ms"%(d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]))
PY
done
if [ "$NONCU" = "" ]; then
CMD="python bench.py --workload msm --logn 22 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:msm_affine_round -s 10 -c 3 -f -o gpurun_out/r02/ncu_affine_round $CMD > gpurun_out/r02/ncu_f.log 2>&1
ncu -i gpurun_out/r02/ncu_affine_round.ncu-rep --page raw --csv > gpurun_out/r02/ncu_affine_round.raw.csv 2>/dev/null
ncu -i gpurun_out/r02/ncu_affine_round.ncu-rep --page source --csv > gpurun_out/r02/ncu_affine_round.source.csv 2>/dev/null
rm -f gpurun_out/r02/*.ncu-rep
fi

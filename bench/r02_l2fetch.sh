#!/bin/bash
mkdir -p gpurun_out/r02
for G in 128 64 32; do
  for W in "msm 24" "msm 22" "ntt 24" "prove 20"; do
    set -- $W
    ZKP_L2_FETCH=$G python bench.py --workload $1 --logn $2 --steps 4 --warmup 3 --no-cpu-baseline --no-prove16 > gpurun_out/r02/l2f${G}_$1$2.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/r02/l2f${G}_$1$2.json')); print('L2fetch $G $1 2^$2: %.3f ms (kernel group %.3f)'%(d['ms_per_step'], d['roofline']['kernel_ms']))"
  done
done

#!/bin/bash
# launch list + one full capture of the batched-affine round kernel (first and second round), msm 2^22
mkdir -p gpurun_out/r02
CMD="python bench.py --workload msm --logn 22 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r02/ncu_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_msm22_aff.csv $CMD > gpurun_out/r02/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_affine_round -s 12 -c 2 -f -o gpurun_out/r02/ncu_affine_round $CMD > gpurun_out/r02/ncu_f.log 2>&1
ncu -i gpurun_out/r02/ncu_affine_round.ncu-rep --page raw --csv > gpurun_out/r02/ncu_affine_round.raw.csv 2>/dev/null
ncu -i gpurun_out/r02/ncu_affine_round.ncu-rep --page source --csv > gpurun_out/r02/ncu_affine_round.source.csv 2>/dev/null
rm -f gpurun_out/r02/*.ncu-rep
ls -la gpurun_out/r02 | tail

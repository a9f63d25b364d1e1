#!/bin/bash
# refresh of the ncu captures behind profiles/traffic.json after a kernel change (plain runs first)
mkdir -p gpurun_out/r02
P20="python bench.py --workload prove --logn 20 --steps 1 --warmup 3 --no-cpu-baseline --no-prove16"
M24="python bench.py --workload msm --logn 24 --steps 1 --warmup 3 --no-cpu-baseline"
$P20 > gpurun_out/r02/plain_prove20.json 2> gpurun_out/r02/plain_prove20.err || exit 1
$M24 > gpurun_out/r02/plain_msm24.json 2> gpurun_out/r02/plain_msm24.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02/launches_prove20.csv $P20 > gpurun_out/r02/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_msm24.csv $M24 > gpurun_out/r02/ncu_l2.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"msm_affine|msm_accumulate" -s 72 -c 8 -f -o gpurun_out/r02/ncu_acc_prove20 $P20 > gpurun_out/r02/ncu_f1.log 2>&1
$NCU -k regex:"msm_affine|msm_accumulate" -s 45 -c 10 -f -o gpurun_out/r02/ncu_acc_msm24 $M24 > gpurun_out/r02/ncu_f2.log 2>&1
for r in gpurun_out/r02/*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f gpurun_out/r02/*.ncu-rep
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02/final_bench_n1.json 2> gpurun_out/r02/final_bench_n1.err; echo "bench rc=$?"

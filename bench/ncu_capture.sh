# One ncu --set full capture per hot kernel (run under gpurun on ONE GPU, after the plain runs passed).
set -x
NCU="ncu --set full --clock-control none --import-source on"
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_sharding.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workload ntt --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ntt_v6.json 2> gpurun_out/bench_ntt_v6.err; cat gpurun_out/bench_ntt_v6.json
python bench.py --workload msm --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_msm_v6.json 2> gpurun_out/bench_msm_v6.err; cat gpurun_out/bench_msm_v6.json
$NCU -k regex:msm_accumulate_kernel -s 3 -c 1 -f -o gpurun_out/ncu_msm_accumulate_v2 python bench.py --workload msm --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_msm_v2.log 2>&1
$NCU -k regex:ntt_pass_kernel -s 9 -c 3 -f -o gpurun_out/ncu_ntt_pass_v2 python bench.py --workload ntt --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ntt_v2.log 2>&1
$NCU -k regex:quotient_kernel -s 3 -c 1 -f -o gpurun_out/ncu_quotient_v1 python bench.py --logn 18 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_quot_v1.log 2>&1
$NCU -k regex:msm_accumulate_kernel -s 15 -c 1 -f -o gpurun_out/ncu_msm_accumulate_prove16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_msm_p16.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_prove_v6.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_prove6.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_msm_v6.csv python bench.py --workload msm --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_msm6.log 2>&1
for r in gpurun_out/*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
# keep the pull small: source-level pages only for the MSM kernel, reports themselves stay on the box
ncu -i gpurun_out/ncu_msm_accumulate_v2.ncu-rep --page source --csv > gpurun_out/ncu_msm_accumulate_v2.source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -20

set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_final.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_prove16_ref.json 2> gpurun_out/final_prove16_ref.err
python bench.py --steps 10 --warmup 3 > gpurun_out/final_prove16.json 2> gpurun_out/final_prove16.err; cut -c1-300 gpurun_out/final_prove16.json
python bench.py --logn 20 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_prove20.json 2> gpurun_out/final_prove20.err
for l in 16 18 20 22 24; do python bench.py --workload msm --logn $l --steps 5 --warmup 3 > gpurun_out/final_msm$l.json 2> gpurun_out/final_msm$l.err; done
for l in 16 18 20 22 24 26; do python bench.py --workload ntt --logn $l --steps 5 --warmup 3 > gpurun_out/final_ntt$l.json 2> gpurun_out/final_ntt$l.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_prove16_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_p16.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_msm22_final.csv python bench.py --workload msm --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_m22.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ntt24_final.csv python bench.py --workload ntt --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_n24.log 2>&1

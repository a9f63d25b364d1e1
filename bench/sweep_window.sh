set -x
python -m pytest tests/test_gpu_msm.py tests/test_gpu_prover.py -x -q -m gpu 2>&1 | tail -3
for logn in 20 22; do for c in 16 17 18 19 20; do
  ZKP_MSM_WINDOW=$c python bench.py --workload msm --logn $logn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_msm${logn}_c$c.json 2> gpurun_out/bench_msm${logn}_c$c.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_msm${logn}_c$c.json').read()); print('WIN logn $logn c $c ms', round(d['ms_per_step'],3), 'acc', round(d['roofline']['kernel_ms'],3))"
done; done
for c in 16 18 19; do
  ZKP_MSM_WINDOW=$c python bench.py --logn 20 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prove20_c$c.json 2> gpurun_out/bench_prove20_c$c.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_prove20_c$c.json').read()); print('WIN prove20 c $c ms', d['ms_per_step'], {k:round(v['ms_per_proof'],2) for k,v in d['kernel_groups'].items()})"
done
for c in 13 14 16; do
  ZKP_MSM_WINDOW=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prove16_c$c.json 2> gpurun_out/bench_prove16_c$c.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_prove16_c$c.json').read()); print('WIN prove16 c $c ms', d['ms_per_step'], {k:round(v['ms_per_proof'],2) for k,v in d['kernel_groups'].items()})"
done

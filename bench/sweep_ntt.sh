set -x
for mb in 2 3 4; do
  ZKP_NTT_BLOCKS_PER_SM=$mb python bench.py --workload ntt --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ntt_occ$mb.json 2> gpurun_out/bench_ntt_occ$mb.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_ntt_occ$mb.json').read()); print('NTTOCC $mb ntt24 ms', d['ms_per_step'])"
  ZKP_NTT_BLOCKS_PER_SM=$mb python bench.py --workload ntt --logn 19 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ntt19_occ$mb.json 2> gpurun_out/bench_ntt19_occ$mb.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_ntt19_occ$mb.json').read()); print('NTTOCC $mb ntt19 ms', d['ms_per_step'])"
done
python bench.py --logn 20 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prove20_v7.json 2> gpurun_out/bench_prove20_v7.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_prove20_v7.json').read()); print('prove20 ms', d['ms_per_step'], {k:round(v['ms_per_proof'],2) for k,v in d['kernel_groups'].items()})"
python bench.py --workload msm --logn 24 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_msm24_v7.json 2> gpurun_out/bench_msm24_v7.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_msm24_v7.json').read()); print('msm24 ms', d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_msm24_v7.csv python bench.py --workload msm --logn 24 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_msm24.log 2>&1

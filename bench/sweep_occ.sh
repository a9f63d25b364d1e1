set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --workload ntt --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ntt_v7.json 2> gpurun_out/bench_ntt_v7.err; cat gpurun_out/bench_ntt_v7.json | cut -c1-400
for mb in 2 3 4 5 6; do
  ZKP_MSM_BLOCKS_PER_SM=$mb python bench.py --workload msm --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_msm_occ$mb.json 2> gpurun_out/bench_msm_occ$mb.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_msm_occ$mb.json').read()); print('OCC $mb msm22 ms', d['ms_per_step'], 'acc', d['roofline']['kernel_ms'])"
  ZKP_MSM_BLOCKS_PER_SM=$mb python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prove_occ$mb.json 2> gpurun_out/bench_prove_occ$mb.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_prove_occ$mb.json').read()); print('OCC $mb prove16 ms', d['ms_per_step'], {k:round(v['ms_per_proof'],2) for k,v in d['kernel_groups'].items()})"
done

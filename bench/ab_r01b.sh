# HISTORICAL (round 1): A/B of the cooperative bucket reduction (ZKP_MSM_REDUCE=legacy) and of software-pipelined
# loads in the accumulate kernel (ZKP_MSM_BLOCKS_PER_SM=7).  Both knobs were removed after the measurement;
# results: profiles/r01/ab_*.json.
set -x
timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prover.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_coop_v1.log
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - gpurun_out/$name.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    g=d.get('kernel_groups',{})
    print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), {k:round(v['ms_per_proof'],3) for k,v in g.items() if v['ms_per_proof']>0}, 'rf', round(d['roofline']['frac'],3), d['roofline'].get('kernel_ms'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
P
}
run ab_prove16_coop python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ZKP_MSM_REDUCE=legacy run ab_prove16_legacy python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ZKP_MSM_BLOCKS_PER_SM=7 run ab_prove16_pf python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run ab_msm22 python bench.py --workload msm --logn 22 --steps 5 --warmup 3 --no-cpu-baseline
ZKP_MSM_BLOCKS_PER_SM=7 run ab_msm22_pf python bench.py --workload msm --logn 22 --steps 5 --warmup 3 --no-cpu-baseline
run ab_prove20 python bench.py --logn 20 --steps 3 --warmup 3 --no-cpu-baseline
ZKP_MSM_BLOCKS_PER_SM=7 run ab_prove20_pf python bench.py --logn 20 --steps 3 --warmup 3 --no-cpu-baseline

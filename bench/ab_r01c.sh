# HISTORICAL (round 1): sweep of the lanes per row/column sum (ZKP_MSM_ROWCOL_LPO, knob removed afterwards;
# the launcher now picks it from the launch size).  Results: profiles/r01/ab2_*.json.
set -x
timeout 300 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prover.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_coop_v2.log
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - gpurun_out/$name.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    g=d.get('kernel_groups',{})
    print(sys.argv[1], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), {k:round(v['ms_per_proof'],3) for k,v in g.items() if v['ms_per_proof']>0}, 'rf', round(d['roofline']['frac'],3), d['roofline'].get('kernel_ms'))
except Exception as e: print(sys.argv[1], 'FAILED', e)
P
}
run ab2_prove16 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ZKP_MSM_ROWCOL_LPO=32 run ab2_prove16_lpo32 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ZKP_MSM_ROWCOL_LPO=16 run ab2_prove16_lpo16 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ZKP_MSM_ROWCOL_LPO=8 run ab2_prove16_lpo8 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ZKP_MSM_ROWCOL_LPO=4 run ab2_prove16_lpo4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run ab2_msm16 python bench.py --workload msm --logn 16 --steps 10 --warmup 3 --no-cpu-baseline
run ab2_msm20 python bench.py --workload msm --logn 20 --steps 5 --warmup 3 --no-cpu-baseline
run ab2_msm22 python bench.py --workload msm --logn 22 --steps 5 --warmup 3 --no-cpu-baseline

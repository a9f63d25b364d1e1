#!/bin/bash
# round 2: single-GPU sweeps (BASELINE configs 3, 4), compile lines (a13), and the ncu refresh behind profiles/traffic.json
mkdir -p gpurun_out/r02
for L in 16 18 20 22 24; do
  python bench.py --workload msm --logn $L --steps 5 --warmup 3 > gpurun_out/r02/sweep_msm${L}_n1.json 2>/dev/null
done
for L in 16 18 20 22 24 26; do
  python bench.py --workload ntt --logn $L --steps 5 --warmup 3 > gpurun_out/r02/sweep_ntt${L}_n1.json 2>/dev/null
done
for L in 16 20; do
  python bench.py --workload compile --logn $L --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02/compile${L}_n1.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/sweep_*_n1.json'))+sorted(glob.glob('gpurun_out/r02/compile*_n1.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], '%.3f ms'%d['ms_per_step'], '%.1f %s'%(d['value'], d['unit']), 'e2e %.1f'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['clocks'])
    except Exception as e: print(f, 'FAILED', e)
PY
P20="python bench.py --workload prove --logn 20 --steps 1 --warmup 3 --no-cpu-baseline --no-prove16"
M24="python bench.py --workload msm --logn 24 --steps 1 --warmup 3 --no-cpu-baseline"
$P20 > gpurun_out/r02/plain_prove20.json 2> gpurun_out/r02/plain_prove20.err || exit 1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"msm_affine|msm_accumulate" -s 72 -c 8 -f -o gpurun_out/r02/ncu_acc_prove20 $P20 > gpurun_out/r02/ncu_f1.log 2>&1
$NCU -k regex:"msm_affine|msm_accumulate" -s 45 -c 10 -f -o gpurun_out/r02/ncu_acc_msm24 $M24 > gpurun_out/r02/ncu_f2.log 2>&1
for r in gpurun_out/r02/*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f gpurun_out/r02/*.ncu-rep

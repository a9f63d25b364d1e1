# Final measurement of round 1 (second half): tests, every bench line, launch lists, one ncu capture.
set -x
O=gpurun_out/final2
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu_final.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/final_prove16_ref.json 2> $O/final_prove16_ref.err
python bench.py --steps 10 --warmup 3 > $O/final_prove16.json 2> $O/final_prove16.err; cut -c1-300 $O/final_prove16.json
python bench.py --logn 20 --steps 3 --warmup 3 --no-cpu-baseline > $O/final_prove20.json 2> $O/final_prove20.err
for l in 16 18 20 22 24; do python bench.py --workload msm --logn $l --steps 5 --warmup 3 > $O/final_msm$l.json 2> $O/final_msm$l.err; done
for l in 16 18 20 22 24 26; do python bench.py --workload ntt --logn $l --steps 5 --warmup 3 > $O/final_ntt$l.json 2> $O/final_ntt$l.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_prove16_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_p16.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_msm22_final.csv python bench.py --workload msm --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_m22.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:msm_rowcol_coop_kernel -s 12 -c 1 -f -o $O/ncu_msm_rowcol_coop python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_rowcol.log 2>&1
$NCU -k regex:msm_accumulate_kernel -s 12 -c 1 -f -o $O/ncu_msm_accumulate_prove16_v2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_acc16.log 2>&1
for r in $O/*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f $O/*.ncu-rep
ls -la $O | tail -40

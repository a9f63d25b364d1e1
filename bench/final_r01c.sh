# Refresh of the proof lines after the last changes of round 1 (evaluation batching, side-stream scheduling).
set -x
O=gpurun_out/final3
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu_final.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/final_prove16_ref.json 2> $O/final_prove16_ref.err
python bench.py --steps 10 --warmup 3 > $O/final_prove16.json 2> $O/final_prove16.err; cut -c1-300 $O/final_prove16.json
python bench.py --logn 20 --steps 3 --warmup 3 --no-cpu-baseline > $O/final_prove20.json 2> $O/final_prove20.err
python bench.py --logn 18 --steps 5 --warmup 3 --no-cpu-baseline > $O/final_prove18.json 2> $O/final_prove18.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_prove16_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_p16.log 2>&1
ls -la $O

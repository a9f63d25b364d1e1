#!/bin/bash
# round 2: four-step NTT (fused twiddle-transpose, native all-to-all): world-1 parity incl. 2^24 / 2^26, then with
# >= 2 GPUs the NCCL worker (incl. the 2^24 sizes) and the strong-scaling NTT / MSM lines
mkdir -p gpurun_out/r02
NG=$(nvidia-smi -L | wc -l)
(timeout 1200 python -m pytest tests/test_gpu_sharding.py -x -q -k "four_step" 2>&1 | tail -5) > gpurun_out/r02/pytest_fourstep_n$NG.log
tail -2 gpurun_out/r02/pytest_fourstep_n$NG.log
if [ $NG -ge 2 ]; then
  ZKP_WORKER_BIG=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29561 tests/multigpu_worker.py > gpurun_out/r02/multigpu_worker_n$NG.log 2>&1
  grep "MULTIGPU\|Error\|assert" gpurun_out/r02/multigpu_worker_n$NG.log | head -5
  for W in "ntt 26" "ntt 24" "msm 24"; do
    set -- $W
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29562 \
      bench.py --gpus $NG --workload $1 --logn $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02/shard_$1$2_n$NG.json 2> gpurun_out/r02/shard_$1$2_n$NG.err
    python - <<PY
import json
t=[l for l in open("gpurun_out/r02/shard_$1$2_n$NG.json") if l.startswith("{")]
if t:
    d=json.loads(t[-1]); print("$1 2^$2 on $NG GPUs: %.3f ms  e2e %.1f %s"%(d["ms_per_step"], d["e2e"]["value"], d["e2e"]["unit"]))
else:
    print("$1 $2 FAILED"); print(open("gpurun_out/r02/shard_$1$2_n$NG.err").read()[-1500:])
PY
  done
fi

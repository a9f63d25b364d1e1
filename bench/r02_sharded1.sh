#!/bin/bash
# round 2: native multi-GPU driver -- world-size-1 tests of the sharded code path (1 GPU), then, with >= 2 GPUs,
# the NCCL worker and the strong-scaling bench
mkdir -p gpurun_out/r02
NG=$(nvidia-smi -L | wc -l)
(timeout 1200 python -m pytest tests/test_gpu_sharding.py -x -q 2>&1 | tail -15) > gpurun_out/r02/pytest_sharding_n$NG.log
tail -5 gpurun_out/r02/pytest_sharding_n$NG.log
if [ $NG -ge 2 ]; then
  for L in 16 20; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $NG --logn $L --steps 5 --warmup 3 --no-cpu-baseline --no-prove16 > gpurun_out/r02/shard_prove${L}_n$NG.json 2> gpurun_out/r02/shard_prove${L}_n$NG.err
    tail -c 1200 gpurun_out/r02/shard_prove${L}_n$NG.json; tail -3 gpurun_out/r02/shard_prove${L}_n$NG.err
  done
fi

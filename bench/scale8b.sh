set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; tail -c 400 gpurun_out/$name.json; tail -2 gpurun_out/$name.err; }
run scale2_prove16_weak_n8 $TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline
for N in 4 8; do run scale2_prove20_shard_n$N $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --shard --logn 20 --steps 3 --warmup 3 --no-cpu-baseline; done
run scale2_msm24_shard_n8 $TR --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --shard --workload msm --logn 24 --steps 3 --warmup 3 --no-cpu-baseline
run scale2_ntt26_shard_n8 $TR --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --shard --workload ntt --logn 26 --steps 3 --warmup 3 --no-cpu-baseline
timeout 600 $TR --nproc-per-node 8 --master-port 29549 tests/multigpu_worker.py > gpurun_out/multigpu_worker_n8_v2.log 2>&1; tail -2 gpurun_out/multigpu_worker_n8_v2.log

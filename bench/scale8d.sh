set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; tail -c 300 gpurun_out/$name.json; tail -2 gpurun_out/$name.err; }
run scale4_prove20_shard_n8 $TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --shard --logn 20 --steps 5 --warmup 3 --no-cpu-baseline
timeout 400 $TR --nproc-per-node 8 --master-port 29549 tests/multigpu_worker.py > gpurun_out/multigpu_worker_n8_v4.log 2>&1; tail -2 gpurun_out/multigpu_worker_n8_v4.log

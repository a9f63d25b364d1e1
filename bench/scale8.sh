set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; tail -c 600 gpurun_out/$name.json; tail -2 gpurun_out/$name.err; }
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/smi8.txt
nvidia-smi topo -m >> gpurun_out/smi8.txt 2>&1
free -g >> gpurun_out/smi8.txt
for N in 2 4 8; do run scale_prove16_weak_n$N $TR --nproc-per-node $N --master-port 2950$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline; done
run scale_msm24_n1 python bench.py --workload msm --logn 24 --steps 3 --warmup 3 --no-cpu-baseline
run scale_ntt26_n1 python bench.py --workload ntt --logn 26 --steps 3 --warmup 3 --no-cpu-baseline
for N in 4 8; do
  run scale_msm24_shard_n$N $TR --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --shard --workload msm --logn 24 --steps 3 --warmup 3 --no-cpu-baseline
  run scale_ntt26_shard_n$N $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --shard --workload ntt --logn 26 --steps 3 --warmup 3 --no-cpu-baseline
  run scale_prove20_shard_n$N $TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --shard --logn 20 --steps 3 --warmup 3 --no-cpu-baseline
done
timeout 600 $TR --nproc-per-node 8 --master-port 29549 tests/multigpu_worker.py > gpurun_out/multigpu_worker_n8.log 2>&1; tail -3 gpurun_out/multigpu_worker_n8.log

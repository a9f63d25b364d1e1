"""diagnostic: repeated sharded proofs, field-by-field differences against the single-GPU proof"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import dusk_plonk_b200 as z
from host_mirror.composer import synthetic_circuit
from host_mirror.synthetic import SplitMix64
from dusk_plonk_b200.field import fr_to_mont1
from dusk_plonk_b200.plonk_params import PlonkParams, ShardedNativeParams
from dusk_plonk_b200.prover import COMM_NAMES, EVAL_NAMES
k = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ctx = z.Context(local)
circ = synthetic_circuit(k)
rng = SplitMix64(8349); tau = rng.fr(); bl = [rng.fr() for _ in range(11)]
ref_prover = z.PlonkKey.compile(PlonkParams.setup_synthetic(ctx, k, fr_to_mont1(tau)), circ)
want, _ = ref_prover.create_proof(bl, circ)
comm = z.NativeComm.from_torch_distributed(ctx)
prover = z.PlonkKey.compile(ShardedNativeParams.setup_synthetic(ctx, comm, k, fr_to_mont1(tau)), circ)
for it in range(6):
    got, _ = prover.create_proof(bl, circ)
    bad = [c for c in COMM_NAMES if getattr(got, c) != getattr(want, c)] + \
          [e for e in EVAL_NAMES if got.evaluations[e] != want.evaluations[e]]
    print("rank %d iter %d differing: %s" % (rank, it, bad), flush=True)
dist.barrier()

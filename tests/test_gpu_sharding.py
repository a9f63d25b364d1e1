"""GPU: the multi-GPU paths -- sharded KZG commit and the four-step NTT -- through the
C ABI against the oracle.  With one visible GPU they run at world_size 1 (same code path, trivial
exchange); with >= 2 GPUs the 2-rank NCCL run of tests/multigpu_worker.py is checked as well."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import cport, curve
from oracle.fields import R_MOD, fr_to_mont_limbs, g1_from_mont_limbs
from oracle.rng import SplitMix64, random_fr_raw_limbs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("k", [4, 9, 13, 16])
@pytest.mark.parametrize("inverse,coset", [(False, False), (True, False), (False, True), (True, True)])
def test_four_step_world1(ctx, cport, k, inverse, coset):
    from dusk_plonk_b200.sharding import FourStepNtt, LocalCommunicator
    n = 1 << k
    host = random_fr_raw_limbs(k * 7 + inverse + 2 * coset, n)
    fs = FourStepNtt(ctx, LocalCommunicator(), k)
    fs.scatter_input(host)
    fs.run(inverse=inverse, coset=coset)
    out = np.zeros((n, 4), dtype=np.uint64)
    fs.gather_output(out)
    assert np.array_equal(out, cport.ntt(host, k, inverse=inverse, coset=coset))


def test_sharded_commit_world1(ctx):
    from dusk_plonk_b200.sharding import LocalCommunicator, ShardedPlonkParams
    from dusk_plonk_b200.plonk_params import PlonkParams
    rng = SplitMix64(3)
    tau = rng.fr()
    taum = fr_to_mont_limbs([tau])[0]
    sp = ShardedPlonkParams.setup_synthetic(ctx, LocalCommunicator(), 9, taum)
    pp = PlonkParams.setup_synthetic(ctx, 9, taum)
    buf = ctx.upload(random_fr_raw_limbs(5, 515))
    assert sp.commit(buf) == pp.commit(buf)


def test_two_gpu_nccl_worker():
    """2 ranks over NCCL: sharded commit + sharded create_proof + four-step NTT (all three kinds)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2); world_size-1 path is covered above")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU OK" in r.stdout

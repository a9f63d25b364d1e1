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


# ---------------------------------------------------------------- native multi-GPU path (zkp_comm, csrc/comm.cu)
@pytest.mark.parametrize("k", [3, 5, 10, 11, 14])
def test_coset8_ntt_is_the_8n_coset_dft(ctx, cport, k):
    """zkp_coset8_ntt_dev on all eight cosets, re-interleaved (point 8 m + u = coset u, element m), equals the
    reference's 8n-point coset_dft -- for n, n + 3 (blinded: folds back) and short inputs."""
    n = 1 << k
    for len_in in (n + 3, n, 2, 2 * n):
        host = random_fr_raw_limbs(100 * k + len_in % 97, len_in)
        src = ctx.upload(host)
        dst = ctx.alloc(8 * n)
        ctx.coset8_ntt(src, 0, len_in, dst, 0, k, 0, 8)
        got = dst.download().reshape(8, n, 4).transpose(1, 0, 2).reshape(8 * n, 4)
        assert np.array_equal(got, cport.ntt(host, k + 3, coset=True)), (k, len_in)
        # a sub-range of cosets is the matching rows
        part = ctx.alloc(3 * n)
        ctx.coset8_ntt(src, 0, len_in, part, 0, k, 4, 3)
        assert np.array_equal(part.download(), dst.download()[4 * n:7 * n])


@pytest.mark.parametrize("name", ["range", "readme", "logic", "chain60", "synthetic12", "synthetic16"])
def test_native_sharded_prover_world1_bit_exact(ctx, name):
    """The multi-GPU code path of the native driver (commit through zkp_commit_batch_sharded_dev, the quotient on
    cosets, coset-wise inverse + combination) on ONE rank: key commitments and the 1040 proof bytes equal the
    single-GPU driver's -- which the prover tests pin to the oracle -- and the committed oracle digests."""
    import hashlib
    import json
    import circuits
    import dusk_plonk_b200 as z
    from host_mirror.composer import SynthesizedCircuit, synthetic_circuit
    from dusk_plonk_b200.field import fr_to_mont1
    from dusk_plonk_b200.plonk_params import PlonkParams, ShardedNativeParams
    rng = SplitMix64(8349)
    tau = rng.fr()
    if name.startswith("synthetic"):
        k = int(name[9:])
        circ, label = synthetic_circuit(k), b"plonk"
    else:
        # chain60: m = 60 gates in n = 64, so trim keeps 2 n + 7 powers -- the rounds-4/5 path that gathers t(X)
        comp = {"range": lambda: circuits.range_circuit(424242), "readme": circuits.readme_circuit,
                "logic": circuits.logic_curve_circuit, "chain60": lambda: circuits.arithmetic_chain(60)}[name]()
        circ, label = SynthesizedCircuit.from_composer(comp), b"demo"
        k = circ.n.bit_length() - 1
    bl = [rng.fr() for _ in range(11)]
    taum = fr_to_mont1(tau)
    comm = z.NativeComm(ctx)
    assert (comm.rank, comm.world) == (0, 1)
    sp = ShardedNativeParams.setup_synthetic(ctx, comm, k + 1, taum)
    sprover = z.PlonkKey.compile_with_circuit(sp, label, circ)
    pprover = z.PlonkKey.compile_with_circuit(PlonkParams.setup_synthetic(ctx, k + 1, taum), label, circ)
    assert dict(sprover.verifier_key) == dict(pprover.verifier_key)
    sproof, spi = sprover.create_proof(bl, circ)
    pproof, ppi = pprover.create_proof(bl, circ)
    assert spi == ppi and sproof.wire_bytes == pproof.wire_bytes
    if name.startswith("synthetic"):
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "synthetic_proofs.json")))
        # (the committed digests are of SRS 2^k + 7; this test's SRS is longer, so prove against that one too)
        sp2 = ShardedNativeParams.setup_synthetic(ctx, comm, k, taum)
        pr2 = z.PlonkKey.compile(sp2, circ)
        proof2, _ = pr2.create_proof(bl, circ)
        assert hashlib.sha256(proof2.wire_bytes).hexdigest() == gold[str(k)]["sha256"]
        pr2.close()
    sprover.close(); pprover.close()
    comm.close()


def test_native_sharded_prover_rejects_unsatisfied_circuit(ctx):
    import circuits
    import dusk_plonk_b200 as z
    from host_mirror.composer import SynthesizedCircuit
    from dusk_plonk_b200.field import fr_to_mont1
    from dusk_plonk_b200.plonk_params import Error, ShardedNativeParams
    rng = SplitMix64(5)
    taum = fr_to_mont1(rng.fr())
    bl = [rng.fr() for _ in range(11)]
    good = SynthesizedCircuit.from_composer(circuits.boolean_select_circuit(bit=1))
    bad = SynthesizedCircuit.from_composer(circuits.boolean_select_circuit(bit=2))
    comm = z.NativeComm(ctx)
    prover = z.PlonkKey.compile_with_circuit(ShardedNativeParams.setup_synthetic(ctx, comm, 6, taum), b"demo", good)
    prover.create_proof(bl, good)
    with pytest.raises(Error):
        prover.create_proof(bl, bad)
    prover.close()
    comm.close()


# ---------------------------------------------------------------- BASELINE sizes of the sharded paths
def _sparse_poly(k, seed):
    n = 1 << k
    idxs = [0, 1, 5, n // 3, n // 2 + 1, n - 1]
    rng = SplitMix64(seed)
    vals = [rng.fr() for _ in idxs]
    sp = np.zeros((n, 4), dtype=np.uint64)
    sp[idxs] = fr_to_mont_limbs(vals)
    return sp, idxs, vals


@pytest.mark.parametrize("inverse,coset", [(False, False), (True, True)])
def test_four_step_2p24_equals_single_gpu_kernel(ctx, inverse, coset):
    """BASELINE config 4 threshold size: the four-step path (column transforms, twiddle, transpose, row
    transforms) against the single-GPU kernel on the same 2^24 random vector, element for element."""
    from dusk_plonk_b200.sharding import FourStepNtt, LocalCommunicator
    k, n = 24, 1 << 24
    host = random_fr_raw_limbs(2400 + inverse, n)
    fs = FourStepNtt(ctx, LocalCommunicator(), k)
    fs.scatter_input(host)
    fs.run(inverse=inverse, coset=coset)
    out = np.zeros((n, 4), dtype=np.uint64)
    fs.gather_output(out)
    ref = ctx.upload(host)
    ctx.ntt_dev(ref, n, ref, k, inverse, coset)
    assert np.array_equal(out, ref.download())


def test_four_step_2p26_spot_horner(ctx):
    """Largest BASELINE size: out[j] = sum_i a_i w^(i j) at sampled j for a sparse polynomial (cheap exact
    Horner on the host), plus the round trip through the inverse four-step transform of a random vector."""
    from dusk_plonk_b200.sharding import FourStepNtt, LocalCommunicator
    from oracle.fields import domain_generator, fr_from_mont_limbs
    k, n = 26, 1 << 26
    sp, idxs, vals = _sparse_poly(k, 26)
    fs = FourStepNtt(ctx, LocalCommunicator(), k)
    fs.scatter_input(sp)
    fs.run()
    out = np.zeros((n, 4), dtype=np.uint64)
    fs.gather_output(out)
    w = domain_generator(k)
    for j in [0, 1, 2, n // 2, n - 1, 123457]:
        x = pow(w, j, R_MOD)
        exp = sum(v * pow(x, i, R_MOD) for i, v in zip(idxs, vals)) % R_MOD
        assert fr_from_mont_limbs(out[j:j + 1]) == [exp], j
    fs.scatter_input(out)
    fs.run(inverse=True)
    back = np.zeros((n, 4), dtype=np.uint64)
    fs.gather_output(back)
    assert np.array_equal(back, sp)


def _known_dlog_commit(tau, sc_mont):
    from oracle.fields import FR_MONT_RINV, _from_limbs_fast
    acc, t = 0, 1
    for v in _from_limbs_fast(sc_mont, 4):
        acc = (acc + v * t) % R_MOD
        t = t * tau % R_MOD
    return curve.mul(curve.G1_GEN, acc * FR_MONT_RINV % R_MOD)


def test_sharded_msm_2p24_known_dlog(ctx):
    """BASELINE config 3 at its largest size through the sharded commit path (SRS range + partial-sum
    combination; one rank here, 2 .. 8 in tests/multigpu_worker.py): (sum_i s_i tau^i) G by known dlog."""
    from dusk_plonk_b200.sharding import LocalCommunicator, ShardedPlonkParams
    n = 1 << 24
    tau = SplitMix64(4242).fr()
    sp = ShardedPlonkParams(ctx, LocalCommunicator(), n, 0, n, ctx.srs_generate(fr_to_mont_limbs([tau])[0], n))
    sc = random_fr_raw_limbs(55555, n)
    got = sp.commit(ctx.upload(sc)).affine()
    assert got == _known_dlog_commit(tau, sc)


def test_native_sharded_commit_2p22_known_dlog(ctx):
    """zkp_commit_batch_sharded_dev (offsets into the SRS, XYZZ partial-sum gather) at 2^22, batch of two, with a
    coefficient beyond the SRS reported as the degree error on the right polynomial only."""
    import dusk_plonk_b200 as z
    from dusk_plonk_b200.plonk_params import Error, ShardedNativeParams
    n = 1 << 22
    tau = SplitMix64(777).fr()
    comm = z.NativeComm(ctx)
    sp = ShardedNativeParams(ctx, ctx.srs_generate(fr_to_mont_limbs([tau])[0], n), comm)
    sc1, sc2 = random_fr_raw_limbs(1, n), random_fr_raw_limbs(2, n - 12345)
    c1, c2 = sp.commit_batch([ctx.upload(sc1), ctx.upload(sc2)])
    assert c1.affine() == _known_dlog_commit(tau, sc1) and c2.affine() == _known_dlog_commit(tau, sc2)
    over = np.zeros((n + 3, 4), dtype=np.uint64)
    over[:n] = sc1
    over[n + 2] = fr_to_mont_limbs([1])[0]
    with pytest.raises(Error):
        sp.commit(ctx.upload(over))
    over[n + 2] = 0
    assert sp.commit(ctx.upload(over)).affine() == c1.affine()       # zeros beyond the SRS are fine
    comm.close()

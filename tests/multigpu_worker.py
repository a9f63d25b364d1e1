"""torchrun worker (one rank per GPU, NCCL): multi-GPU parity checks against the oracle.
Launched by tests/test_gpu_sharding.py or by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_worker.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import dusk_plonk_b200 as z
    from dusk_plonk_b200.sharding import Communicator, FourStepNtt, ShardedPlonkParams
    from dusk_plonk_b200.plonk_params import PlonkParams
    from host_mirror.composer import SynthesizedCircuit
    from oracle import cport, plonk as oplonk
    from oracle.fields import fr_to_mont_limbs
    from oracle.rng import SplitMix64, random_fr_raw_limbs
    import circuits

    ctx = z.Context(local)
    comm = Communicator(torch.device("cuda", local))
    # 1. four-step NTT, every kind, natural order in / out
    ncomm_ntt = z.NativeComm.from_torch_distributed(ctx) if world in (2, 4, 8) else None
    for k, use_native in ((10, False), (10, True), (15, True), (20, True)):
        if use_native and ncomm_ntt is None:
            continue
        n = 1 << k
        host = random_fr_raw_limbs(k, n)
        # torch.distributed all-to-all (tensors) and the library's own NCCL all-to-all (ffi.NativeComm)
        fs = FourStepNtt(ctx, ncomm_ntt if use_native else comm, k)
        for inverse, coset in ((False, False), (True, False), (False, True), (True, True)):
            fs.scatter_input(host)
            fs.run(inverse=inverse, coset=coset)
            out = np.zeros((n, 4), dtype=np.uint64)
            fs.gather_output(out)
            t = torch.from_numpy(out.view(np.int64)).cuda()
            dist.all_reduce(t)                                   # disjoint row slabs: sum = union
            got = t.cpu().numpy().view(np.uint64)
            assert np.array_equal(got, cport.ntt(host, k, inverse=inverse, coset=coset)), (k, inverse, coset)
    # 2. sharded commit and a whole proof with sharded commits: bit-identical to one GPU / the oracle
    rng = SplitMix64(8349)
    tau = rng.fr()
    taum = fr_to_mont_limbs([tau])[0]
    cs = circuits.readme_circuit()
    circ = SynthesizedCircuit.from_composer(cs)
    k = circ.n.bit_length() - 1
    sp = ShardedPlonkParams.setup_synthetic(ctx, comm, k + 1, taum)
    prover = z.PlonkKey.compile_with_circuit(sp, b"demo", circ)
    commit = oplonk.default_commit(tau=tau)
    opk, ovk = oplonk.compile_circuit(circ, commit, sp.total_len)
    from oracle.merlin import Transcript as OTranscript
    otr = OTranscript.base(b"demo", oplonk.vk_transcript_list(ovk), circ.m)
    bl = [rng.fr() for _ in range(11)]
    gproof, gpi = prover.create_proof(bl, circ)
    oproof, opi = oplonk.create_proof(opk, circ, commit, otr, bl)
    for c in oplonk.Proof.COMM_NAMES:
        assert getattr(gproof, c) == getattr(oproof, c), c
    assert gproof.evaluations == oproof.evaluations
    assert oplonk.verify(ovk, circ.n, gproof, circ.pi_indexes, gpi, otr, oplonk.trapdoor_kzg_check(tau))
    # 3. the native multi-GPU driver (zkp_comm: NCCL inside libzkp_b200.so, no Python between the rounds): the README
    #    circuit against the oracle proof above, then the synthetic 2^16-gate circuit bench.py times against the
    #    committed oracle digest -- byte-identical on every rank
    import hashlib
    import json
    from host_mirror.composer import synthetic_circuit
    from dusk_plonk_b200.plonk_params import ShardedNativeParams
    if world in (1, 2, 4, 8):
        ncomm = z.NativeComm.from_torch_distributed(ctx) if world > 1 else z.NativeComm(ctx)
        np_ = ShardedNativeParams.setup_synthetic(ctx, ncomm, k + 1, taum)
        nprover = z.PlonkKey.compile_with_circuit(np_, b"demo", circ)
        nproof, npi = nprover.create_proof(bl, circ)
        assert nproof.wire_bytes == oproof.to_bytes(), "native sharded proof differs from the oracle"
        nprover.close()
        # a tiny circuit (n = 32): with 8 ranks a slab holds 4 coefficients, fewer than t_4's tail -- the gather path
        rc_ = SynthesizedCircuit.from_composer(circuits.range_circuit(99))
        rk = rc_.n.bit_length() - 1
        rprover = z.PlonkKey.compile_with_circuit(ShardedNativeParams.setup_synthetic(ctx, ncomm, rk + 1, taum), b"demo", rc_)
        ropk, rovk = oplonk.compile_circuit(rc_, commit, (1 << (rk + 1)) + 7)
        rotr = OTranscript.base(b"demo", oplonk.vk_transcript_list(rovk), rc_.m)
        rproof, _ = rprover.create_proof(bl, rc_)
        roproof, _ = oplonk.create_proof(ropk, rc_, commit, rotr, bl)
        assert rproof.wire_bytes == roproof.to_bytes(), "native sharded proof of the range circuit differs from the oracle"
        rprover.close()
        circ16 = synthetic_circuit(16)
        rng16 = SplitMix64(8349)
        tau16 = fr_to_mont_limbs([rng16.fr()])[0]
        bl16 = [rng16.fr() for _ in range(11)]
        p16 = z.PlonkKey.compile(ShardedNativeParams.setup_synthetic(ctx, ncomm, 16, tau16), circ16)
        proof16, _ = p16.create_proof(bl16, circ16)
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "synthetic_proofs.json")))
        dg = hashlib.sha256(proof16.wire_bytes).hexdigest()
        assert dg == gold["16"]["sha256"], "2^16 sharded proof differs from the committed oracle digest"
        digs = [None] * world
        dist.all_gather_object(digs, dg)
        assert all(d == dg for d in digs)
        colls, sent = ncomm.stats()
        assert world == 1 or (colls > 0 and sent > 0)
        p16.close()
        ncomm.close()
    # 4. BASELINE sizes (set ZKP_WORKER_BIG=1: minutes of host-side checking): four-step NTT 2^24 against the
    #    single-GPU kernel run on this rank's own GPU, sharded MSM 2^24 by known discrete log
    if os.environ.get("ZKP_WORKER_BIG"):
        from oracle import curve
        from oracle.fields import FR_MONT_RINV, R_MOD, _from_limbs_fast, g1_from_mont_limbs
        k24, n24 = 24, 1 << 24
        host = random_fr_raw_limbs(2424, n24)
        fs = FourStepNtt(ctx, ncomm_ntt if ncomm_ntt is not None else comm, k24)
        for inverse, coset in ((False, False), (True, True)):
            fs.scatter_input(host)
            fs.run(inverse=inverse, coset=coset)
            mine = np.zeros((n24, 4), dtype=np.uint64)
            fs.gather_output(mine)
            ref = ctx.upload(host)
            ctx.ntt_dev(ref, n24, ref, k24, inverse, coset)
            full = ref.download().reshape(fs.C, fs.R, 4)[:, rank * fs.Rl:(rank + 1) * fs.Rl, :]
            got = mine.reshape(fs.C, fs.R, 4)[:, rank * fs.Rl:(rank + 1) * fs.Rl, :]
            assert np.array_equal(got, full), ("four-step 2^24", inverse, coset)
        del fs
        tau24 = SplitMix64(4242).fr()
        lo, hi = z.sharding.shard_range(n24, rank, world)
        sp24 = ShardedPlonkParams(ctx, comm, n24, lo, hi, ctx.srs_generate(fr_to_mont_limbs([tau24])[0], hi - lo, first=lo))
        sc = random_fr_raw_limbs(55555, n24)
        got = sp24.commit(ctx.upload(sc)).affine()
        if rank == 0:
            acc, t = 0, 1
            for v in _from_limbs_fast(sc, 4):
                acc = (acc + v * t) % R_MOD
                t = t * tau24 % R_MOD
            assert got == curve.mul(curve.G1_GEN, acc * FR_MONT_RINV % R_MOD), "sharded MSM 2^24"
        digs = [None] * world
        dist.all_gather_object(digs, got)
        assert all(d == got for d in digs)
    dist.barrier()
    if rank == 0:
        print("MULTIGPU OK world=%d%s" % (world, " (+2^24 sizes)" if os.environ.get("ZKP_WORKER_BIG") else ""), flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

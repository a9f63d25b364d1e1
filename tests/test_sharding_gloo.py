"""CPU, world_size 2 over gloo: the host-side logic of the sharded KZG commit
(dusk-plonk_b200/sharding.py).  The device MSM is replaced by an oracle-backed
stand-in so that only the range split, the 96-byte all-gather and the partial-sum
combination are under test; the GPU parity of the MSM itself is tests/test_gpu_msm.py."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeBuf:
    def __init__(self, vals):
        self.vals, self.n = list(vals), len(vals)


class FakeSrs:
    def __init__(self, tau, first, n):
        self.tau, self.first, self.n = tau, first, n


class FakeCtx:
    """poly_degree / msm_dev / srs_generate with the CPU oracle behind them."""

    def srs_generate(self, tau, n, first=0):
        return FakeSrs(tau, first, n)

    def poly_degree(self, buf, off, n):
        nz = [i for i in range(n) if buf.vals[off + i]]
        return nz[-1] if nz else -1

    def ref(self, buf, off=0, n=None):
        return (buf, off, n)

    def commit_batch_dev(self, srs, refs):
        from oracle import curve
        from oracle.fields import R_MOD, g1_to_mont_limbs
        out = []
        for buf, off, n in refs:
            assert n <= srs.n
            acc = 0
            for i in range(n):
                acc = (acc + buf.vals[off + i] * pow(srs.tau, srs.first + i, R_MOD)) % R_MOD
            out.append(g1_to_mont_limbs([curve.mul(curve.G1_GEN, acc)])[0])
        return np.stack(out), [0] * len(refs)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dusk_plonk_b200.plonk_params import Error
        from dusk_plonk_b200.sharding import Communicator, ShardedPlonkParams, shard_range
        from oracle import curve
        from oracle.fields import R_MOD
        from oracle.ntt import poly_eval
        from oracle.rng import SplitMix64
        comm = Communicator()
        assert (comm.rank, comm.world) == (rank, world)
        blobs = comm.all_gather_bytes(bytes([rank]) * 96)
        assert blobs == [bytes([r]) * 96 for r in range(world)]
        rng = SplitMix64(8349)
        tau = rng.fr()
        pp = ShardedPlonkParams.setup_synthetic(FakeCtx(), comm, 5, tau)     # 39 powers
        assert (pp.lo, pp.hi) == shard_range(39, rank, world) and pp.srs.first == pp.lo
        out = {}
        for name, coeffs in (("dense", [rng.fr() for _ in range(37)]),
                             ("low_half_only", [rng.fr() for _ in range(10)] + [0] * 25),
                             ("trailing_zeros", [rng.fr() for _ in range(39)] + [0] * 100),
                             ("zero", [0] * 20)):
            got = pp.commit(FakeBuf(coeffs)).affine()
            exp = curve.mul(curve.G1_GEN, poly_eval(coeffs, tau)) if any(coeffs) else None
            assert got == exp, name
            out[name] = got
        with pytest.raises(Error):
            pp.commit(FakeBuf([1] * 40))
        assert pp.commit_or_default(FakeBuf([1] * 40)).is_identity()
        small = pp.trim(8)                                                   # keeps 15 powers
        assert small.total_len == 15 and (small.lo, small.hi) == shard_range(15, rank, world)
        c = [rng.fr() for _ in range(15)]
        assert small.commit(FakeBuf(c)).affine() == curve.mul(curve.G1_GEN, poly_eval(c, tau))
        q.put((rank, "ok", out["dense"]))
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_sharded_commit_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), [r[2] for r in res if r[1] != "ok"]
    assert res[0][2] == res[1][2]     # every rank derives the same commitment -> same transcript


def _slice_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dusk_plonk_b200.sharding import Communicator
        comm = Communicator()
        for n in (64, 1024):
            per = n // world
            # element i of rank r's vector: limbs (r, i, i * i, 7)
            i = torch.arange(n, dtype=torch.int64)
            full = torch.stack([torch.full_like(i, rank), i, i * i, torch.full_like(i, 7)], dim=1).reshape(-1).clone()
            send = torch.empty(world * (per + 8) * 4, dtype=torch.int64)
            recv = torch.empty(world * (per + 8) * 4, dtype=torch.int64)
            comm.exchange_slices(recv, full, send, n)
            got = recv.view(world, per + 8, 4)
            idx = (rank * per + torch.arange(per + 8, dtype=torch.int64)) % n     # my slice + the 8-element halo
            for src in range(world):
                want = torch.stack([torch.full_like(idx, src), idx, idx * idx, torch.full_like(idx, 7)], dim=1)
                assert torch.equal(got[src], want), (n, src)
        q.put((rank, "ok", None))
    except Exception:
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slice_exchange_gloo(world):
    """The all-to-all of quotient slices of the sharded proof (Communicator.exchange_slices): rank s receives
    elements [s per, (s + 1) per + 8) mod n of every rank's vector -- the halo wraps around for the last rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slice_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), [r[2] for r in res if r[1] != "ok"]


def test_shard_ranges_cover_exactly():
    from dusk_plonk_b200.sharding import shard_range
    for total in (0, 1, 7, 39, 65543):
        for world in (1, 2, 3, 8):
            r = [shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1

"""Multi-GPU KZG commit: SRS powers split into per-GPU index ranges.

One process per GPU.  Rank r keeps the powers [tau^i]_1 for i in [lo_r, hi_r) -- and their
window table -- resident; a commit is a local MSM over the matching coefficient range, an
all-gather of one 96-byte affine point per rank and G - 1 host additions.  The polynomial
itself is replicated (every rank runs the cheap NTT / element-wise rounds), so no bulk data
crosses NVLink: the exchange is latency only.  ``ShardedPlonkParams`` is a drop-in for
``PlonkParams`` (same ``commit`` / ``commit_or_default`` / ``trim``), so ``PlonkKey.compile``
and ``Prover.create_proof`` run unchanged and every rank derives the same transcript.
"""
import numpy as np

from .ffi import BufferView
from .field import g1_add, g1_from_bytes, g1_from_mont, g1_to_bytes
from .plonk_params import Error
from .poly_commit import Commitment


class Communicator:
    """All-gather of small fixed-size byte blobs over ``torch.distributed`` (NCCL on GPUs,
    gloo in the CPU tests)."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")

    def all_gather_bytes(self, blob: bytes):
        t = self.torch.frombuffer(bytearray(blob), dtype=self.torch.uint8).to(self.device)
        out = self.torch.empty(self.world * len(blob), dtype=self.torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        raw = out.cpu().numpy().tobytes()
        return [raw[i * len(blob):(i + 1) * len(blob)] for i in range(self.world)]


class LocalCommunicator:
    """world_size 1 stand-in (and the unit-test double)."""
    rank, world = 0, 1

    def all_gather_bytes(self, blob):
        return [blob]


def shard_range(total, rank, world):
    """Contiguous, balanced index range of rank in [0, total)."""
    return total * rank // world, total * (rank + 1) // world


def combine_partials(blobs):
    acc = None
    for b in blobs:
        acc = g1_add(acc, g1_from_bytes(b))
    return acc


class ShardedPlonkParams:
    def __init__(self, ctx, comm, total_len, lo, hi, srs, tau_mont=None):
        self.ctx, self.comm = ctx, comm
        self.total_len, self.lo, self.hi = total_len, lo, hi
        self.srs = srs          # this rank's powers [lo, hi) with their window table
        self._tau = tau_mont

    @classmethod
    def setup_synthetic(cls, ctx, comm, k, tau_mont):
        total = (1 << k) + 7
        lo, hi = shard_range(total, comm.rank, comm.world)
        return cls(ctx, comm, total, lo, hi, ctx.srs_generate(tau_mont, hi - lo, first=lo), tau_mont)

    @classmethod
    def from_points(cls, ctx, comm, xy):
        xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, 12)
        lo, hi = shard_range(xy.shape[0], comm.rank, comm.world)
        p = cls(ctx, comm, xy.shape[0], lo, hi, ctx.srs_load(xy[lo:hi]))
        p._host_points = xy
        return p

    def max_degree(self):
        return self.total_len - 1

    def trim(self, n):
        keep = min(self.total_len, n + 7)
        if keep == self.total_len:
            return self
        lo, hi = shard_range(keep, self.comm.rank, self.comm.world)
        if self._tau is not None:
            srs = self.ctx.srs_generate(self._tau, hi - lo, first=lo)
            return ShardedPlonkParams(self.ctx, self.comm, keep, lo, hi, srs, self._tau)
        p = ShardedPlonkParams(self.ctx, self.comm, keep, lo, hi, self.ctx.srs_load(self._host_points[lo:hi]))
        p._host_points = self._host_points[:keep]
        return p

    def commit(self, poly):
        buf, off, n = (poly.buf, poly.off, poly.n) if isinstance(poly, BufferView) else (poly, 0, poly.n)
        top = self.ctx.poly_degree(buf, off, n)
        if top >= self.total_len:
            raise Error("polynomial degree exceeds the SRS")   # identical decision on every rank
        a, b = self.lo, min(self.hi, top + 1)
        part = None
        if b > a:
            part = g1_from_mont(self.ctx.msm_dev(self.srs, buf, off + a, b - a))
        return Commitment.from_affine(combine_partials(self.comm.all_gather_bytes(g1_to_bytes(part))))

    def commit_batch(self, polys):
        return [self.commit(p) for p in polys]

    def commit_or_default(self, poly):
        try:
            return self.commit(poly)
        except Error:
            return Commitment(np.zeros(12, dtype=np.uint64))

"""Multi-GPU KZG commit: SRS powers split into per-GPU index ranges.

One process per GPU.  Rank r keeps the powers [tau^i]_1 for i in [lo_r, hi_r) -- and their
window table -- resident; a commit is a local MSM over the matching coefficient range, an
all-gather of one 96-byte affine point per rank and G - 1 host additions.  The polynomial
itself is replicated (every rank runs the cheap NTT / element-wise rounds), so no bulk data
crosses NVLink: the exchange is latency only.  ``ShardedPlonkParams`` is a drop-in for
``PlonkParams`` (same ``commit`` / ``commit_or_default`` / ``trim``), so ``PlonkKey.compile``
and ``Prover.create_proof`` run unchanged and every rank derives the same transcript.
"""
import ctypes

import numpy as np

from .ffi import BufferView
from .field import g1_add, g1_from_bytes, g1_from_mont, g1_to_bytes
from .plonk_params import Error
from .poly_commit import Commitment


class Communicator:
    """All-gather over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests): small byte blobs
    (partial commitments) and device-resident slabs (coset evaluations, quotient slices)."""

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


    def all_gather_device(self, out, mine, stage):
        """``out`` (world equal chunks, device tensor) <- every rank's ``mine`` (which may be a view of
        ``out``; it is staged through ``stage`` so the collective never aliases its input)."""
        st = stage[:mine.numel()]
        st.copy_(mine)
        self.dist.all_gather_into_tensor(out, st)
        self.torch.cuda.synchronize()


    def exchange_slices(self, recv, full, send, n):
        """All-to-all of slices: this rank holds ``full`` (n Fr as 4 int64 each); rank s receives its
        elements [s per, (s + 1) per + 8) (indices mod n, per = n / world) from every rank.  ``recv``
        ends up as [source rank][per + 8] Fr."""
        torch, G = self.torch, self.world
        per = n // G
        f4 = full[:n * 4].view(n, 4)
        s3 = send[:G * (per + 8) * 4].view(G, per + 8, 4)
        s3[:, :per] = f4.view(G, per, 4)
        starts = ((torch.arange(G, device=full.device) + 1) * per) % n
        s3[:, per:] = f4[starts[:, None] + torch.arange(8, device=full.device)[None, :]]
        self.dist.all_to_all_single(recv, send[:G * (per + 8) * 4])
        if full.is_cuda:
            torch.cuda.synchronize()


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
    def __init__(self, ctx, comm, total_len, lo, hi, srs=None, tau_mont=None):
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
        return self.commit_batch([poly])[0]

    def commit_batch(self, polys):
        """Local batched MSM over each polynomial's [lo, hi) coefficients, ONE all-gather of
        len(polys) x 96 bytes, G - 1 host additions per commitment."""
        refs, views = [], []
        for p in polys:
            buf, off, n = (p.buf, p.off, p.n) if isinstance(p, BufferView) else (p, 0, p.n)
            views.append((buf, off, n))
            if n > self.total_len and self.ctx.poly_degree(buf, off + self.total_len, n - self.total_len) >= 0:
                raise Error("polynomial degree exceeds the SRS")   # identical decision on every rank
            cnt = max(0, min(self.hi, n) - self.lo)
            refs.append(self.ctx.ref(buf, off + (self.lo if cnt else 0), cnt))
        out, _ = self.ctx.commit_batch_dev(self.srs, refs)
        blob = b"".join(g1_to_bytes(g1_from_mont(out[i])) for i in range(len(polys)))
        parts = self.comm.all_gather_bytes(blob)
        return [Commitment.from_affine(combine_partials([pb[96 * i:96 * (i + 1)] for pb in parts]))
                for i in range(len(polys))]

    def commit_or_default(self, poly):
        try:
            return self.commit(poly)
        except Error:
            return Commitment(np.zeros(12, dtype=np.uint64))


# ====================================================================== four-step NTT
class FourStepNtt:
    """Fr NTT of 2^k points over G GPUs with one all-to-all: N = R x C.

    With i = r C + c and j = jc R + jr,
        X[jc R + jr] = sum_c w_C^(c jc) * w_N^(c jr) * sum_r x[r C + c] w_R^(r jr).
    Rank s owns the columns c in [s C/G, (s+1) C/G) on input and the rows jr in
    [s R/G, (s+1) R/G) on output:

        in_buf   [c_loc][r]     this rank's columns, each contiguous   (``scatter_input`` lays it out)
        step 1   C/G size-R NTTs over r (batched single-GPU kernels)   -> [c_loc][jr]
        step 2   * w_N^(c jr), transpose to [jr][c_loc]
        step 3   all-to-all: rows jr of rank t's slab go to rank t     (NCCL over NVLink)
        step 4   regroup to [jr_loc][c], R/G size-C NTTs over c        -> out_buf [jr_loc][jc]

    so ``out_buf[jr_loc][jc] = X[jc R + jr]``.  The reference-facing call hands over / receives host
    vectors in natural order (``Fft::dft`` by value), and host <-> device copies can place elements
    anywhere, so the column / row slabs cost nothing extra there (``scatter_input`` /
    ``gather_output``).  Inverse and coset variants follow ``Fft::{idft, coset_dft, coset_idft}``:
    1/N falls out of the two inverse sub-transforms, the coset factors g^i / g^-j are separable
    (g^(r C + c) = (g^C)^r g^c) and applied by ``zkp_scale_matrix_dev``.
    """

    def __init__(self, ctx, comm, k, torch_device=None):
        import torch
        self.torch = torch
        self.ctx, self.comm, self.k = ctx, comm, k
        G = comm.world
        self.kr = k - k // 2            # R = 2^kr rows, C = 2^kc columns
        self.kc = k // 2
        self.R, self.C = 1 << self.kr, 1 << self.kc
        assert self.C % G == 0 and self.R % G == 0, "world size must divide both factors"
        self.Cl, self.Rl = self.C // G, self.R // G
        self.nloc = (1 << k) // G
        # two ping-pong slabs.  With the library's own communicator (``ffi.NativeComm``: NCCL inside
        # libzkp_b200.so, the all-to-all is queued on the context's stream, no host synchronisation) or on one
        # rank they are plain device vectors; with the torch communicator NCCL addresses them as tensors
        self.native = hasattr(comm, "h") or not hasattr(comm, "dist")
        if self.native:
            self.a, self.b = ctx.alloc(self.nloc), ctx.alloc(self.nloc)
        else:
            dev = torch_device if torch_device is not None else torch.device("cuda", ctx.device)
            self.ta = torch.empty(self.nloc * 4, dtype=torch.int64, device=dev)
            self.tb = torch.empty(self.nloc * 4, dtype=torch.int64, device=dev)
            self.a = ctx.wrap(self.ta.data_ptr(), self.nloc)
            self.b = ctx.wrap(self.tb.data_ptr(), self.nloc)

    # ---- host <-> device placement (natural-order host vector)
    def scatter_input(self, host):
        """host: (N, 4) uint64 natural order -> in_buf [c_loc][r] of this rank: one strided copy of the rank's
        column slab (R rows of Cl elements, pitch C) and a transpose on the device -- the host never reorders."""
        s = self.comm.rank
        host = np.ascontiguousarray(host, dtype=np.uint64).reshape(self.R * self.C, 4)
        base = host.ctypes.data + s * self.Cl * 32
        bb, bo = _base_of(self.b)
        self.ctx.check(self.ctx.lib.zkp_buf_upload_2d(self.ctx.h, bb.h, bo, ctypes.c_void_p(base), self.Cl, self.R, self.C))
        self.ctx.permute(self.b, 0, self.a, 0, self.Cl, self.R, 1)       # [r][c_loc] -> [c_loc][r]

    def gather_output(self, host_out):
        """out_buf [jr_loc][jc] -> host_out[jc R + jr] for this rank's rows (other ranks' stay): transpose on the
        device, one strided copy."""
        s = self.comm.rank
        assert host_out.flags["C_CONTIGUOUS"] and host_out.dtype == np.uint64
        self.ctx.permute(self.a, 0, self.b, 0, self.C, self.Rl, 1)       # [jr_loc][jc] -> [jc][jr_loc]
        base = host_out.ctypes.data + s * self.Rl * 32
        bb, bo = _base_of(self.b)
        self.ctx.check(self.ctx.lib.zkp_buf_download_2d(self.ctx.h, bb.h, bo, ctypes.c_void_p(base), self.Rl, self.C, self.R))

    # ---- the transform: in_buf (self.a) -> out_buf (self.a)
    def run(self, inverse=False, coset=False):
        from .ffi import fft_constant
        ctx, G, s = self.ctx, self.comm.world, self.comm.rank
        R, C, Rl, Cl, k = self.R, self.C, self.Rl, self.Cl, self.k
        a, b = self.a, self.b
        g, gi = fft_constant(k, 3), fft_constant(k, 4)
        if coset and not inverse:
            # x[r C + c] *= g^(r C + c): rows a = c_loc, columns b = r
            ctx.scale_matrix(a, 0, Cl, R, s * Cl, g, _pow_mont(ctx, g, C), 1)
        # step 1: size-R transforms of the Cl local columns
        ctx.ntt_dev_batch(a, R, R, a, R, self.kr, inverse, False, Cl)
        # step 2: twiddle w_N^(c jr) fused into the transpose that makes a destination's rows contiguous:
        # one pass over the data, two-level twiddle table (zkp_twiddle_transpose_dev)
        ctx.twiddle_transpose(a, 0, b, 0, Cl, R, s * Cl, k, inverse)      # b[jr][c_loc]
        # step 3: all-to-all (rank t receives its rows jr from every source)
        if G > 1:
            if self.native:
                self.comm.all_to_all(b, a, Rl * Cl)          # on the context's stream, behind the transpose
            else:
                ctx.sync()
                self.comm.dist.all_to_all_single(self.ta, self.tb)
                self.torch.cuda.synchronize()
            src = a                                      # [src][jr_loc][c_loc]
            dst = b
        else:
            src, dst = b, a
        # step 4: regroup to [jr_loc][c = src * Cl + c_loc] and transform over c
        if G > 1:
            ctx.permute(src, 0, dst, 0, Rl, G, Cl)       # dst[jr_loc][src][c_loc]
            work = dst
        else:
            work = src
        ctx.ntt_dev_batch(work, C, C, a, C, self.kc, inverse, False, Rl)
        if coset and inverse:
            # X[jc R + jr] *= g^-(jc R + jr): rows a = jr_loc, columns b = jc
            ctx.scale_matrix(a, 0, Rl, C, s * Rl, gi, _pow_mont(ctx, gi, R), 1)
        ctx.sync()


def _base_of(buf):
    from .ffi import _base
    return _base(buf)


def _pow_mont(ctx, base_mont, e):
    from .field import fr_from_mont, fr_to_mont1, R_MOD
    return fr_to_mont1(pow(fr_from_mont([base_mont])[0], e, R_MOD))

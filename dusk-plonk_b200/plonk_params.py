"""Host-side mirror of ``zksnarks::plonk::PlonkParams`` for the commit path
(``keypair.commit(&poly)?``, src/prover.rs:133; ``.unwrap_or_default()``, src/key.rs:138;
``keypair.trim(additional_n)``, src/key.rs:82), backed by the CUDA MSM."""
import numpy as np

from .ffi import BufferView, Context, DeviceBuffer, ZkpError, ZKP_ERR_DEGREE
from .poly_commit import Coefficients, Commitment


class Error(Exception):
    """Stands in for ``zksnarks::error::Error`` on the commit path."""


class PlonkParams:
    def __init__(self, ctx: Context, srs, opening_key=None):
        self.ctx = ctx
        self.srs = srs
        self.opening_key = opening_key   # EvaluationKey: the G2 side ([tau]_2), carried along by trim

    @classmethod
    def setup_synthetic(cls, ctx, k, tau_mont):
        """SRS with the structure of ``PlonkParams::setup(k, rng)`` (tests/range.rs:26):
        [tau^i]_1 for i < 2^k + 7 (the prover needs n + 7 powers for t_4).
        tau is an explicit input because the reference's RNG derivation is unreachable."""
        from .verifier import EvaluationKey
        return cls(ctx, ctx.srs_generate(tau_mont, (1 << k) + 7), EvaluationKey.from_tau(tau_mont))

    @classmethod
    def from_points(cls, ctx, xy, beta_h=None):
        """An SRS handed over as points (a ceremony's output): G1 powers and, for verification, [tau]_2."""
        from .verifier import EvaluationKey
        return cls(ctx, ctx.srs_load(xy), EvaluationKey(beta_h) if beta_h is not None else None)

    def verification_key(self):
        """``keypair.verification_key()`` (src/key.rs:320): the ``EvaluationKey`` a ``Verifier`` opens with."""
        if self.opening_key is None:
            raise Error("this SRS was loaded without its G2 element [tau]_2")
        return self.opening_key

    def max_degree(self):
        return self.srs.n - 1

    def trim(self, n):
        """Keep n + 7 powers (``trim`` must leave room for the blinded quotient chunk)."""
        keep = min(self.srs.n, n + 7)
        if keep == self.srs.n:
            return self
        return PlonkParams(self.ctx, self.srs.trim(keep), self.opening_key)   # device-side slice: no download / re-upload

    def commit(self, poly):
        """-> Commitment, raising ``Error`` when degree > SRS (Err in the reference)."""
        try:
            if isinstance(poly, DeviceBuffer):
                return Commitment(self.ctx.commit_dev(self.srs, poly))
            if isinstance(poly, BufferView):
                return Commitment(self.ctx.commit_dev(self.srs, poly.buf, poly.off, poly.n))
            v = poly.v if isinstance(poly, Coefficients) else poly
            return Commitment(self.ctx.commit(self.srs, v))
        except ZkpError as e:
            if e.code == ZKP_ERR_DEGREE:
                raise Error("polynomial degree exceeds the SRS") from e
            raise

    def commit_batch(self, polys):
        """Several ``commit`` calls whose results are needed together (the four wire polynomials,
        the four quotient chunks) as one batched launch sequence.  Raises ``Error`` if any of
        them fails the degree check, like the first ``?`` in the reference would."""
        refs = []
        for p in polys:
            b, off, n = (p.buf, p.off, p.n) if isinstance(p, BufferView) else (p, 0, p.n)
            refs.append(self.ctx.ref(b, off, n))
        out, status = self.ctx.commit_batch_dev(self.srs, refs)
        if any(st == ZKP_ERR_DEGREE for st in status):
            raise Error("polynomial degree exceeds the SRS")
        return [Commitment(out[i]) for i in range(len(polys))]

    def commit_or_default(self, poly):
        """``keypair.commit(&p).unwrap_or_default()`` (src/key.rs:138-154)."""
        try:
            return self.commit(poly)
        except Error:
            return Commitment(np.zeros(12, dtype=np.uint64))


class ShardedNativeParams(PlonkParams):
    """``PlonkParams`` of one rank of a job shared by several GPUs (BASELINE config 5), all inside
    libzkp_b200.so: every commitment is split by SRS ranges over the ranks of ``comm`` (``zkp_comm``, NCCL) and
    the partial sums are gathered and added on every rank, so all ranks see the same commitment and derive
    the same transcript.  The SRS (and its window table) is complete on every rank: 2 GiB at 2^20 gates.
    ``PlonkKey.compile`` / ``Prover.create_proof`` see an ordinary commit key; a key compiled from it owns
    only this rank's cosets of the 8n domain and proves through ``zkp_prover_create_sharded``."""

    def __init__(self, ctx, srs, comm, opening_key=None):
        super().__init__(ctx, srs, opening_key)
        self.native_comm = comm

    @classmethod
    def setup_synthetic(cls, ctx, comm, k, tau_mont):
        from .verifier import EvaluationKey
        return cls(ctx, ctx.srs_generate(tau_mont, (1 << k) + 7), comm, EvaluationKey.from_tau(tau_mont))

    def trim(self, n):
        keep = min(self.srs.n, n + 7)
        if keep == self.srs.n:
            return self
        return ShardedNativeParams(self.ctx, self.srs.trim(keep), self.native_comm, self.opening_key)

    def commit(self, poly):
        return self.commit_batch([poly])[0]

    def commit_batch(self, polys):
        refs = []
        for p in polys:
            b, off, n = (p.buf, p.off, p.n) if isinstance(p, BufferView) else (p, 0, p.n)
            refs.append(self.ctx.ref(b, off, n))
        out, status = self.ctx.commit_batch_sharded_dev(self.native_comm, self.srs, refs)
        if any(st == ZKP_ERR_DEGREE for st in status):
            raise Error("polynomial degree exceeds the SRS")
        return [Commitment(out[i]) for i in range(len(polys))]

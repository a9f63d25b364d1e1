"""Host-side mirror of the ``poly_commit`` types the reference prover uses
(``use poly_commit::{Coefficients, Fft, PointsValue}``, src/prover.rs:18), backed by the
CUDA library.  Same method names and argument meaning as the Rust API reconstructed in
SURVEY Appendix A, so parity tests read like the reference's own call sites:

    fft = Fft(ctx, k)                      # Fft::<Fr>::new(k)            src/prover.rs:88
    a_w_poly = fft.idft(a_w_scalar)        # fft.idft(PointsValue)        src/prover.rs:121
    evals = fft_8n.coset_dft(z_poly)       # zero-pads short inputs       quotient_poly.rs:54

Vectors are (n, 4) uint64 numpy arrays of Montgomery limbs (the reference's ``Vec<Fr>``
memory image), or ``DeviceBuffer`` handles when they should stay resident in HBM.
"""
import numpy as np

from .ffi import Context, DeviceBuffer, as_fr_array, fft_constant


class Coefficients:
    """``poly_commit::Coefficients<Fr>(pub Vec<Fr>)``."""

    def __init__(self, limbs):
        self.v = as_fr_array(limbs)

    def __len__(self):
        return self.v.shape[0]

    def degree(self):
        nz = np.nonzero(self.v.any(axis=1))[0]
        return int(nz[-1]) if len(nz) else 0


class PointsValue:
    """``poly_commit::PointsValue<Fr>(pub Vec<Fr>)``."""

    def __init__(self, limbs):
        self.v = as_fr_array(limbs)

    def __len__(self):
        return self.v.shape[0]


class Commitment:
    """``poly_commit::Commitment<G1Affine>(pub G1Affine)``: 12 uint64 Montgomery limbs
    (x || y), the identity encoded as zeros."""

    def __init__(self, xy):
        self.xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(12)
        self._affine = False

    @classmethod
    def from_affine(cls, pt):
        """From canonical coordinates ((x, y) ints or None)."""
        from .field import P_MOD
        c = cls(np.zeros(12, dtype=np.uint64))
        if pt is not None:
            r = (1 << 384) % P_MOD
            raw = (pt[0] * r % P_MOD).to_bytes(48, "little") + (pt[1] * r % P_MOD).to_bytes(48, "little")
            c.xy = np.frombuffer(raw, dtype="<u8").copy()
        c._affine = pt
        return c

    def affine(self):
        """Canonical (x, y) or None for the identity."""
        if self._affine is False:
            from .field import g1_from_mont
            self._affine = g1_from_mont(self.xy)
        return self._affine

    def is_identity(self):
        return not self.xy.any()

    def __eq__(self, other):
        return isinstance(other, Commitment) and bool((self.xy == other.xy).all())


def _limbs(x):
    return x.v if isinstance(x, (Coefficients, PointsValue)) else x


class Fft:
    """``poly_commit::Fft<Fr>``."""

    def __init__(self, ctx: Context, k: int):
        self.ctx = ctx
        self.k = k
        self._elements = None

    def size(self):
        return 1 << self.k

    def generator(self):
        return fft_constant(self.k, 0)

    def generator_inv(self):
        return fft_constant(self.k, 1)

    def size_inv(self):
        return fft_constant(self.k, 2)

    @property
    def elements(self):
        if self._elements is None:
            self._elements = self.ctx.fft_elements(self.k).download()
        return self._elements

    def _run(self, x, inverse, coset, out_cls):
        if isinstance(x, DeviceBuffer):
            out = x if x.n >= self.size() else self.ctx.alloc(self.size())
            self.ctx.ntt_dev(x, min(x.n, self.size()), out, self.k, inverse, coset)
            return out
        return out_cls(self.ctx.ntt(_limbs(x), self.k, inverse, coset))

    def dft(self, coeffs):
        return self._run(coeffs, False, False, PointsValue)

    def idft(self, evals):
        return self._run(evals, True, False, Coefficients)

    def coset_dft(self, coeffs):
        return self._run(coeffs, False, True, PointsValue)

    def coset_idft(self, evals):
        return self._run(evals, True, True, Coefficients)

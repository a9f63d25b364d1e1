"""BLS12-381 field constants and Montgomery limb encodings (oracle; test infrastructure).

Layouts follow what the reference pins in-tree: ``BlsScalar([u64; 4])`` is the
little-endian Montgomery representation with R = 2^256 (``src/lib.rs:583-588`` is the
one known-answer: MINUS_ONE).  Fq is taken as 6 x u64 LE Montgomery, R = 2^384 (the
zkcrypto/dusk convention; [EXT-RECALL], the ``bls-12-381`` crate is absent).
"""
import numpy as np

# scalar field Fr (255 bit) and base field Fq (381 bit)
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB

FR_LIMBS64 = 4
FQ_LIMBS64 = 6
FR_BITS = 256
FQ_BITS = 384
FR_MONT_R = (1 << FR_BITS) % R_MOD
FQ_MONT_R = (1 << FQ_BITS) % P_MOD
FR_MONT_R2 = (FR_MONT_R * FR_MONT_R) % R_MOD
FQ_MONT_R2 = (FQ_MONT_R * FQ_MONT_R) % P_MOD
FR_MONT_RINV = pow(FR_MONT_R, -1, R_MOD)
FQ_MONT_RINV = pow(FQ_MONT_R, -1, P_MOD)
# -m^-1 mod 2^32 / 2^64 (Montgomery reduction constants)
FR_INV32 = (-pow(R_MOD, -1, 1 << 32)) % (1 << 32)
FR_INV64 = (-pow(R_MOD, -1, 1 << 64)) % (1 << 64)
FQ_INV32 = (-pow(P_MOD, -1, 1 << 32)) % (1 << 32)
FQ_INV64 = (-pow(P_MOD, -1, 1 << 64)) % (1 << 64)

# FftField constants (zkstd; values are the zkcrypto/dusk ones)
TWO_ADICITY = 32
MULTIPLICATIVE_GENERATOR = 7  # also the coset shift used by coset_dft / coset_idft
ROOT_OF_UNITY = pow(MULTIPLICATIVE_GENERATOR, (R_MOD - 1) >> TWO_ADICITY, R_MOD)

# coset representatives of the permutation argument, src/permutation.rs:28-30
K1, K2, K3 = 7, 13, 17


def fr_inv(a):
    return pow(a, -1, R_MOD)


def fq_inv(a):
    return pow(a, -1, P_MOD)


def domain_generator(k):
    """w_k: primitive 2^k-th root of unity, ``Fft::new(k).generator()``."""
    assert 0 <= k <= TWO_ADICITY
    return pow(ROOT_OF_UNITY, 1 << (TWO_ADICITY - k), R_MOD)


# ---------------------------------------------------------------- limb codecs
def _to_limbs(vals, nlimbs):
    out = np.empty((len(vals), nlimbs), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(vals):
        for j in range(nlimbs):
            out[i, j] = (v >> (64 * j)) & mask
    return out


def _from_limbs(arr, nlimbs):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, nlimbs)
    out = []
    for row in arr:
        v = 0
        for j in range(nlimbs):
            v |= int(row[j]) << (64 * j)
        out.append(v)
    return out


def _to_limbs_fast(vals, nlimbs):
    """Vectorised int -> limb conversion through bytes (for large arrays)."""
    nb = nlimbs * 8
    buf = b"".join(int(v).to_bytes(nb, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(len(vals), nlimbs).copy()


def _from_limbs_fast(arr, nlimbs):
    arr = np.ascontiguousarray(np.asarray(arr, dtype="<u8").reshape(-1, nlimbs))
    nb = nlimbs * 8
    raw = arr.tobytes()
    return [int.from_bytes(raw[i * nb:(i + 1) * nb], "little") for i in range(arr.shape[0])]


def fr_to_mont_limbs(vals):
    """canonical ints -> (n,4) uint64 Montgomery limbs (the reference's Fr layout)."""
    return _to_limbs_fast([(v % R_MOD) * FR_MONT_R % R_MOD for v in vals], FR_LIMBS64)


def fr_from_mont_limbs(arr):
    return [v * FR_MONT_RINV % R_MOD for v in _from_limbs_fast(arr, FR_LIMBS64)]


def fr_to_raw_limbs(vals):
    return _to_limbs_fast(vals, FR_LIMBS64)


def fr_from_raw_limbs(arr):
    return _from_limbs_fast(arr, FR_LIMBS64)


def fq_to_mont_limbs(vals):
    return _to_limbs_fast([(v % P_MOD) * FQ_MONT_R % P_MOD for v in vals], FQ_LIMBS64)


def fq_from_mont_limbs(arr):
    return [v * FQ_MONT_RINV % P_MOD for v in _from_limbs_fast(arr, FQ_LIMBS64)]


def g1_to_mont_limbs(points):
    """list of affine points ((x,y) or None for infinity) -> (n,12) uint64.

    ABI convention of include/zkp_b200.h: infinity is encoded as x = y = 0 (not on
    y^2 = x^3 + 4, so unambiguous)."""
    flat = []
    for pt in points:
        if pt is None:
            flat += [0, 0]
        else:
            flat += [pt[0], pt[1]]
    return fq_to_mont_limbs(flat).reshape(len(points), 12)


def g1_from_mont_limbs(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 12)
    vals = fq_from_mont_limbs(arr.reshape(-1, 6))
    out = []
    for i in range(arr.shape[0]):
        x, y = vals[2 * i], vals[2 * i + 1]
        out.append(None if (x == 0 and y == 0) else (x, y))
    return out

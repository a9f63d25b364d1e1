"""Host-side Fr / Fq / G1 encodings at the C-ABI boundary (Python ints <-> Montgomery limbs).

``BlsScalar([u64; 4])`` is little-endian Montgomery with R = 2^256 (pinned by MINUS_ONE,
src/lib.rs:583-588); Fq is 6 x u64 with R = 2^384; G1 affine is x || y with x = y = 0 for
the identity (include/zkp_b200.h).  Challenges, blinders and evaluations cross the boundary
through these helpers; bulk vectors stay on the device."""
import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_FR_R = (1 << 256) % R_MOD
_FR_RINV = pow(_FR_R, -1, R_MOD)
_FQ_RINV = pow((1 << 384) % P_MOD, -1, P_MOD)
K1, K2, K3 = 7, 13, 17  # src/permutation.rs:28-30


def fr_to_mont(vals):
    """iterable of canonical ints -> (n, 4) uint64 Montgomery limbs."""
    vals = list(vals)
    buf = b"".join(((int(v) % R_MOD) * _FR_R % R_MOD).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(len(vals), 4).copy()


def fr_to_mont1(v):
    return fr_to_mont([v])[0]


def fr_from_mont(arr):
    arr = np.ascontiguousarray(np.asarray(arr, dtype="<u8").reshape(-1, 4))
    raw = arr.tobytes()
    return [int.from_bytes(raw[i * 32:(i + 1) * 32], "little") * _FR_RINV % R_MOD for i in range(arr.shape[0])]


def fr_column_to_mont(col, n):
    """Selector / wire column (list of ints or integer ndarray) -> (n, 4) Montgomery limbs,
    zero padded; repeated values are converted once."""
    out = np.zeros((n, 4), dtype=np.uint64)
    if isinstance(col, np.ndarray) and col.dtype != object:
        uniq, inv = np.unique(col, return_inverse=True)
        table = fr_to_mont([int(u) for u in uniq])
        out[:len(col)] = table[inv]
        return out
    cache = {}
    idx = np.empty(len(col), dtype=np.int64)
    keys = []
    for i, v in enumerate(col):
        j = cache.get(v)
        if j is None:
            j = cache[v] = len(keys)
            keys.append(v)
        idx[i] = j
    if keys:
        out[:len(col)] = fr_to_mont(keys)[idx]
    return out


_FQ_R = (1 << 384) % P_MOD


def g1_to_mont(pt):
    """(x, y) canonical ints / None -> 12 x uint64 Montgomery limbs (zeros for the identity)."""
    if pt is None:
        return np.zeros(12, dtype=np.uint64)
    raw = (pt[0] * _FQ_R % P_MOD).to_bytes(48, "little") + (pt[1] * _FQ_R % P_MOD).to_bytes(48, "little")
    return np.frombuffer(raw, dtype="<u8").copy()


def g1_from_mont(xy):
    """12 x uint64 Montgomery (x, y) -> (x, y) canonical ints, or None for the identity."""
    raw = np.ascontiguousarray(np.asarray(xy, dtype="<u8").reshape(12)).tobytes()
    x = int.from_bytes(raw[:48], "little") * _FQ_RINV % P_MOD
    y = int.from_bytes(raw[48:], "little") * _FQ_RINV % P_MOD
    return None if (x == 0 and y == 0) else (x, y)


# ---- host-side G1 addition (used only to combine per-GPU partial commitments: G - 1 adds)
def g1_add(p, q):
    """Affine (x, y) / None points on y^2 = x^3 + 4."""
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    return (x3, (lam * (x1 - x3) - y1) % P_MOD)


def g1_to_bytes(p):
    """96 raw bytes (x || y big-endian, zeros for the identity): the inter-rank wire format."""
    return bytes(96) if p is None else p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")


def g1_from_bytes(b):
    x, y = int.from_bytes(b[:48], "big"), int.from_bytes(b[48:96], "big")
    return None if x == 0 and y == 0 else (x, y)


def g1_mul(p, k):
    """Host double-and-add (decoding checks only)."""
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def g1_decompress(b, subgroup_check=True):
    """Inverse of ``transcript.g1_compress`` (48-byte zcash-style compressed G1).  Untrusted bytes:
    every malformed encoding raises ``ValueError`` (never ``assert``): wrong length / missing compression
    flag, x >= p, stray bits in the encoding of infinity, x not on the curve, and -- BLS12-381 G1 has a
    cofactor -- points outside the r-torsion ([r]P != O) unless ``subgroup_check`` is False."""
    if len(b) != 48 or not b[0] & 0x80:
        raise ValueError("not a compressed G1 point")
    if b[0] & 0x40:
        if b[0] & 0x3F or any(b[1:]):
            raise ValueError("non-canonical encoding of the point at infinity")
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= P_MOD:
        raise ValueError("non-canonical x coordinate")
    y2 = (pow(x, 3, P_MOD) + 4) % P_MOD
    y = pow(y2, (P_MOD + 1) // 4, P_MOD)          # p = 3 mod 4
    if y * y % P_MOD != y2:
        raise ValueError("x is not on the curve")
    if (y > P_MOD - y) != bool(b[0] & 0x20):
        y = P_MOD - y
    if subgroup_check and g1_mul((x, y), R_MOD) is not None:
        raise ValueError("point is not in the prime-order subgroup")
    return (x, y)

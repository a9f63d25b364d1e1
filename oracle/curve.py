"""BLS12-381 G1 in Python big integers (oracle; test infrastructure).

Restates what ``poly_commit::msm_curve_addition`` / ``PlonkParams::commit`` compute
(call sites ``src/prover.rs:133-136,194,262-265,440,452``, ``src/key.rs:138-159``,
``src/prover/proof.rs:507-526``): sum_i s_i * P_i on y^2 = x^3 + 4 over Fq, returned
as an affine point.  A group element is unique, so any correct algorithm is a valid
oracle; ``msm_naive`` is the definition, ``msm_pippenger`` the bucket method the
reference is believed to use ([EXT-RECALL]).
Affine points are ``(x, y)`` tuples of canonical ints, ``None`` is infinity.
"""
from .fields import P_MOD, R_MOD

B_COEFF = 4
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)

_p = P_MOD


def is_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % _p == 0


def neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % _p)


# Jacobian (X, Y, Z): x = X/Z^2, y = Y/Z^3; Z == 0 is infinity.
J_INF = (1, 1, 0)


def to_jac(pt):
    return J_INF if pt is None else (pt[0], pt[1], 1)


def to_affine(j):
    X, Y, Z = j
    if Z == 0:
        return None
    zi = pow(Z, -1, _p)
    zi2 = zi * zi % _p
    return (X * zi2 % _p, Y * zi2 * zi % _p)


def jdouble(j):
    X, Y, Z = j
    if Z == 0 or Y == 0:
        return J_INF
    A = X * X % _p
    B = Y * Y % _p
    C = B * B % _p
    D = 2 * ((X + B) * (X + B) - A - C) % _p
    E = 3 * A % _p
    F = E * E % _p
    X3 = (F - 2 * D) % _p
    Y3 = (E * (D - X3) - 8 * C) % _p
    Z3 = 2 * Y * Z % _p
    return (X3, Y3, Z3)


def jadd(j1, j2):
    X1, Y1, Z1 = j1
    X2, Y2, Z2 = j2
    if Z1 == 0:
        return j2
    if Z2 == 0:
        return j1
    Z1Z1 = Z1 * Z1 % _p
    Z2Z2 = Z2 * Z2 % _p
    U1 = X1 * Z2Z2 % _p
    U2 = X2 * Z1Z1 % _p
    S1 = Y1 * Z2 * Z2Z2 % _p
    S2 = Y2 * Z1 * Z1Z1 % _p
    if U1 == U2:
        if S1 == S2:
            return jdouble(j1)
        return J_INF
    H = (U2 - U1) % _p
    I = 4 * H * H % _p
    J = H * I % _p
    r = 2 * (S2 - S1) % _p
    V = U1 * I % _p
    X3 = (r * r - J - 2 * V) % _p
    Y3 = (r * (V - X3) - 2 * S1 * J) % _p
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % _p
    return (X3, Y3, Z3)


def jadd_affine(j1, pt):
    if pt is None:
        return j1
    return jadd(j1, (pt[0], pt[1], 1))


def jmul(j, k):
    k %= R_MOD
    acc = J_INF
    for bit in bin(k)[2:] if k else "":
        acc = jdouble(acc)
        if bit == "1":
            acc = jadd(acc, j)
    return acc


def mul(pt, k):
    return to_affine(jmul(to_jac(pt), k))


def add(p1, p2):
    return to_affine(jadd(to_jac(p1), to_jac(p2)))


def batch_to_affine(jacs):
    """Montgomery batch inversion; infinities map to None."""
    zs = [j[2] for j in jacs]
    prefix = []
    acc = 1
    for z in zs:
        prefix.append(acc)
        if z:
            acc = acc * z % _p
    inv = pow(acc, -1, _p)
    out = [None] * len(jacs)
    for i in range(len(jacs) - 1, -1, -1):
        z = zs[i]
        if z == 0:
            continue
        zi = inv * prefix[i] % _p
        inv = inv * z % _p
        zi2 = zi * zi % _p
        out[i] = (jacs[i][0] * zi2 % _p, jacs[i][1] * zi2 * zi % _p)
    return out


def msm_naive(points, scalars):
    """The definition: sum_i scalars[i] * points[i] (double-and-add per term)."""
    acc = J_INF
    for pt, s in zip(points, scalars):
        if pt is None or s % R_MOD == 0:
            continue
        acc = jadd(acc, jmul(to_jac(pt), s))
    return to_affine(acc)


def msm_pippenger(points, scalars, c=None):
    """Bucket method (unsigned windows): the algorithm class the reference uses."""
    n = len(scalars)
    if n == 0:
        return None
    if c is None:
        c = 3 if n < 32 else max(3, n.bit_length() * 69 // 100 + 2)
    nwin = (255 + c - 1) // c
    total = J_INF
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            total = jdouble(total)
        buckets = [J_INF] * ((1 << c) - 1)
        for pt, s in zip(points, scalars):
            d = ((s % R_MOD) >> (w * c)) & ((1 << c) - 1)
            if d and pt is not None:
                buckets[d - 1] = jadd_affine(buckets[d - 1], pt)
        run = J_INF
        acc = J_INF
        for b in reversed(buckets):
            run = jadd(run, b)
            acc = jadd(acc, run)
        total = jadd(total, acc)
    return to_affine(total)


def srs_powers(tau, n):
    """[tau^i]_1 for i < n -- the structure ``PlonkParams::setup`` produces
    (``tests/range.rs:26``).  Uses a fixed-base window table so 2^12 powers take
    seconds in Python."""
    # 8-bit fixed-base table: tbl[w][d] = d * 2^(8w) * G
    nwin = 32
    tbl = []
    base = to_jac(G1_GEN)
    for _ in range(nwin):
        row = [J_INF]
        for d in range(1, 256):
            row.append(jadd(row[-1], base))
        tbl.append(batch_to_affine(row))
        for _ in range(8):
            base = jdouble(base)
    out = []
    t = 1
    for _ in range(n):
        acc = J_INF
        for w in range(nwin):
            d = (t >> (8 * w)) & 255
            if d:
                acc = jadd_affine(acc, tbl[w][d])
        out.append(acc)
        t = t * tau % R_MOD
    return batch_to_affine(out)


def commit_known_dlog(dlogs, scalars):
    """Full-size MSM oracle when bases are P_i = d_i * G with known d_i:
    sum s_i P_i = (sum s_i d_i mod r) * G."""
    acc = 0
    for d, s in zip(dlogs, scalars):
        acc = (acc + d * s) % R_MOD
    return mul(G1_GEN, acc)

"""BLS12-381 pairing in Python big integers (oracle; TEST INFRASTRUCTURE).

The acceptance check of the restated verifier: ``batch_check`` ends in a two-pairing product
(``src/commitment_scheme.rs:52-64``: ``multi_miller_loop([(−W, [tau]_2), (C, [1]_2)]).final_exp()``
must be the identity of Gt).  The ``ec-pairing`` / ``bls-12-381`` crates are absent from the
reference tree; this is the textbook ate pairing on the public BLS12-381 parameters: Fq12 as
Fq[w] / (w^12 - 2 w^6 + 2), G2 on y^2 = x^3 + 4(u + 1) over Fq2 = Fq[u] / (u^2 + 1), Miller loop
over |x| = 0xd201000000010000, final exponentiation (p^12 - 1) / r.  Checked by bilinearity and
non-degeneracy in tests/test_oracle_pairing.py.  Slow (seconds per check) and only ever run by tests.
"""
from .fields import P_MOD, R_MOD

_p = P_MOD
ATE_LOOP_COUNT = 0xD201000000010000
FQ12_MODULUS = (2, 0, 0, 0, 0, 0, -2, 0, 0, 0, 0, 0)   # w^12 = 2 w^6 - 2

# G2 generator (x = x0 + x1 u, y = y0 + y1 u)
G2_GEN = (
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)


# ------------------------------------------------------------------ Fq2 (tuples (a0, a1))
def f2_add(a, b): return ((a[0] + b[0]) % _p, (a[1] + b[1]) % _p)
def f2_sub(a, b): return ((a[0] - b[0]) % _p, (a[1] - b[1]) % _p)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % _p, (a[0] * b[1] + a[1] * b[0]) % _p)
def f2_scalar(a, k): return (a[0] * k % _p, a[1] * k % _p)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, _p)
    return (a[0] * d % _p, (-a[1]) * d % _p)


B2 = (4, 4)


def g2_on_curve(pt):
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), B2)) == (0, 0)


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_scalar(f2_mul(x1, x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(pt, k):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, pt)
        pt = g2_add(pt, pt)
        k >>= 1
    return acc


# ------------------------------------------------------------------ Fq12 (lists of 12 ints)
def f12(c):
    return [x % _p for x in c] + [0] * (12 - len(c))


F12_ONE = f12([1])


def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for i in range(22, 11, -1):   # w^i = 2 w^(i-6) - 2 w^(i-12)
        c = t[i]
        if c:
            t[i - 6] += 2 * c
            t[i - 12] -= 2 * c
    return [x % _p for x in t[:12]]


def f12_add(a, b): return [(x + y) % _p for x, y in zip(a, b)]
def f12_sub(a, b): return [(x - y) % _p for x, y in zip(a, b)]


def _deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _poly_rounded_div(a, b):
    dega, degb = _deg(a), _deg(b)
    temp = list(a)
    o = [0] * len(a)
    inv_lead = pow(b[degb], -1, _p)
    for i in range(dega - degb, -1, -1):
        o[i] = (o[i] + temp[degb + i] * inv_lead) % _p
        for c in range(degb + 1):
            temp[c + i] = (temp[c + i] - o[i] * b[c]) % _p
    return o[:_deg(o) + 1]


def f12_inv(a):
    """Extended Euclid in Fq[w] against the modulus polynomial."""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [c % _p for c in FQ12_MODULUS] + [1]
    while _deg(low):
        r = _poly_rounded_div(high, low)
        r += [0] * (13 - len(r))
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] -= lm[i] * r[j]
                new[i + j] -= low[i] * r[j]
        nm = [x % _p for x in nm]
        new = [x % _p for x in new]
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, _p)
    return [x * inv0 % _p for x in lm[:12]]


def f12_div(a, b): return f12_mul(a, f12_inv(b))


def f12_pow(a, e):
    out = F12_ONE
    while e:
        if e & 1:
            out = f12_mul(out, a)
        a = f12_mul(a, a)
        e >>= 1
    return out


_W = f12([0, 1])
_W2 = f12_mul(_W, _W)
_W3 = f12_mul(_W2, _W)
_W2_INV, _W3_INV = f12_inv(_W2), f12_inv(_W3)


def twist(pt):
    """G2 point over Fq2 -> the isomorphic curve over Fq12 (y^2 = x^3 + 4)."""
    if pt is None:
        return None
    (x0, x1), (y0, y1) = pt
    nx = f12([x0 - x1, 0, 0, 0, 0, 0, x1])
    ny = f12([y0 - y1, 0, 0, 0, 0, 0, y1])
    return (f12_mul(nx, _W2_INV), f12_mul(ny, _W3_INV))


def cast_g1(pt):
    return None if pt is None else (f12([pt[0]]), f12([pt[1]]))


def _linefunc(P1, P2, T):
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if x1 != x2:
        m = f12_div(f12_sub(y2, y1), f12_sub(x2, x1))
    elif y1 == y2:
        m = f12_div(f12_mul(f12([3]), f12_mul(x1, x1)), f12_mul(f12([2]), y1))
    else:
        return f12_sub(xt, x1)
    return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))


def _ec12_double(P):
    x, y = P
    m = f12_div(f12_mul(f12([3]), f12_mul(x, x)), f12_mul(f12([2]), y))
    nx = f12_sub(f12_mul(m, m), f12_mul(f12([2]), x))
    return (nx, f12_sub(f12_mul(m, f12_sub(x, nx)), y))


def _ec12_add(P, Q):
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        return _ec12_double(P) if y1 == y2 else None
    m = f12_div(f12_sub(y2, y1), f12_sub(x2, x1))
    nx = f12_sub(f12_sub(f12_mul(m, m), x1), x2)
    return (nx, f12_sub(f12_mul(m, f12_sub(x1, nx)), y1))


def miller_loop(Q2, P1):
    """Q2: G2 affine over Fq2, P1: G1 affine; without the final exponentiation."""
    if Q2 is None or P1 is None:
        return F12_ONE
    Q, P = twist(Q2), cast_g1(P1)
    R, f = Q, F12_ONE
    for i in range(ATE_LOOP_COUNT.bit_length() - 2, -1, -1):
        f = f12_mul(f12_mul(f, f), _linefunc(R, R, P))
        R = _ec12_double(R)
        if (ATE_LOOP_COUNT >> i) & 1:
            f = f12_mul(f, _linefunc(R, Q, P))
            R = _ec12_add(R, Q)
    return f


def final_exponentiation(f):
    return f12_pow(f, (P_MOD ** 12 - 1) // R_MOD)


def pairing(Q2, P1):
    return final_exponentiation(miller_loop(Q2, P1))


def pairing_product_is_one(pairs):
    """prod_i e(P_i, Q_i) == 1 for pairs (P_i in G1, Q_i in G2): one shared final exponentiation,
    like ``multi_miller_loop(..).final_exp()``."""
    f = F12_ONE
    for P1, Q2 in pairs:
        f = f12_mul(f, miller_loop(Q2, P1))
    return final_exponentiation(f) == F12_ONE


def kzg_pairing_check(tau_h):
    """``batch_check``'s final test with the opening key [tau]_2 = tau_h, [1]_2 = G2_GEN
    (``EvaluationKey { prepared_h, prepared_beta_h }``, src/commitment_scheme.rs:51-60)."""
    def check(total_w_neg, total_c):
        return pairing_product_is_one([(total_w_neg, tau_h), (total_c, G2_GEN)])
    return check

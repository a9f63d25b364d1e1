"""Scalar side of the gate widgets: the coefficients of the linearisation polynomial.

Host arithmetic on the 16 opened evaluations (a few hundred Fr multiplications per proof),
mirroring ``widget.linearize`` / ``permutation.linearize`` as called from
``src/prover/linearization_poly.rs:75-105,136-225``.  The widget bodies live in the absent
``zksnarks`` crate ([EXT-RECALL] upstream dusk-plonk 0.13); the permutation
part is pinned by the in-tree verifier identity (``src/prover/proof.rs:386-440``).  The
device evaluates the same identities point-wise in ``quotient_kernel`` (csrc/prover.cu).
"""
from .field import R_MOD as _r, K1, K2, K3

EDWARDS_D = (-(10240 * pow(10241, -1, _r))) % _r


def _delta(f):
    return f * (f - 1) % _r * (f - 2) % _r * (f - 3) % _r


def range_term(sep, a, b, c, d, d_next):
    k = sep * sep % _r
    k2 = k * k % _r
    return (_delta((c - 4 * d) % _r) + _delta((b - 4 * c) % _r) * k + _delta((a - 4 * b) % _r) * k2
            + _delta((d_next - 4 * a) % _r) * (k2 * k % _r)) % _r * sep % _r


def logic_term(sep, a, a_next, b, b_next, c, d, d_next, q_c):
    k = sep * sep % _r
    k2, k3 = k * k % _r, k * k % _r * k % _r
    k4 = k3 * k % _r
    A, B, D = (a_next - 4 * a) % _r, (b_next - 4 * b) % _r, (d_next - 4 * d) % _r
    f = c * (c * (4 * c - 18 * (A + B) + 81) % _r + 18 * (A * A + B * B) - 81 * (A + B) + 83) % _r
    e = (3 * (A + B + D) - 2 * f) % _r
    bb = q_c * (9 * D - 3 * (A + B)) % _r
    return ((c - A * B) % _r * k3 + _delta(A) + _delta(B) * k + _delta(D) * k2 + (bb + e) * k4) % _r * sep % _r


def fixed_base_term(sep, a, a_next, b, b_next, c, d, d_next, q_l, q_r, q_c):
    k = sep * sep % _r
    k2, k3 = k * k % _r, k * k % _r * k % _r
    bit = (d_next - 2 * d) % _r
    y_alpha = (bit * bit % _r * (q_r - 1) + 1) % _r
    x_alpha = bit * q_l % _r
    t = c * a % _r * b % _r * EDWARDS_D % _r
    x_acc = ((a_next + a_next * t) - (a * y_alpha + b * x_alpha)) % _r * k2
    y_acc = ((b_next - b_next * t) - (b * y_alpha + a * x_alpha)) % _r * k3
    return (bit * (bit - 1) % _r * (bit + 1) + x_acc + y_acc + (bit * q_c - c) % _r * k) % _r * sep % _r


def var_base_term(sep, a, a_next, b, b_next, c, d, d_next):
    k = sep * sep % _r
    y1x2, y1y2, x1x2 = b * c % _r, b * d % _r, a * c % _r
    t = EDWARDS_D * d_next % _r * y1x2 % _r
    x3c = ((d_next + y1x2) - (a_next + a_next * t)) % _r * k
    y3c = ((y1y2 + x1x2) - (b_next - b_next * t)) % _r * (k * k % _r)
    return ((a * d - d_next) + x3c + y3c) % _r * sep % _r


def linearization_scalars(n, ch, ev):
    """[(polynomial name, scalar)] with r(X) = sum scalar * poly(X)."""
    alpha, beta, gamma, rs, ls, fs, vs, zc = ch
    a, b, c, d = ev["a_eval"], ev["b_eval"], ev["c_eval"], ev["d_eval"]
    an, bn, dn = ev["a_next_eval"], ev["b_next_eval"], ev["d_next_eval"]
    qa, qc, ql, qr = ev["q_arith_eval"], ev["q_c_eval"], ev["q_l_eval"], ev["q_r_eval"]
    s1, s2, s3, pe = ev["s_sigma_1_eval"], ev["s_sigma_2_eval"], ev["s_sigma_3_eval"], ev["perm_eval"]
    out = [("q_m", a * b % _r * qa % _r), ("q_l", a * qa % _r), ("q_r", b * qa % _r), ("q_o", c * qa % _r),
           ("q_d", d * qa % _r), ("q_c", qa),
           ("q_range", range_term(rs, a, b, c, d, dn)),
           ("q_logic", logic_term(ls, a, an, b, bn, c, d, dn, qc)),
           ("q_fixed_group_add", fixed_base_term(fs, a, an, b, bn, c, d, dn, ql, qr, qc)),
           ("q_variable_group_add", var_base_term(vs, a, an, b, bn, c, d, dn))]
    zh = (pow(zc, n, _r) - 1) % _r
    l1 = zh * pow(n * (zc - 1) % _r, -1, _r) % _r
    x = (a + beta * zc + gamma) % _r * ((b + beta * K1 % _r * zc + gamma) % _r) % _r \
        * ((c + beta * K2 % _r * zc + gamma) % _r) % _r * ((d + beta * K3 % _r * zc + gamma) % _r) % _r * alpha % _r
    y = (-((a + beta * s1 + gamma) % _r * ((b + beta * s2 + gamma) % _r) % _r * ((c + beta * s3 + gamma) % _r) % _r
           * beta % _r * pe % _r * alpha)) % _r
    out.append(("z", (x + l1 * alpha % _r * alpha) % _r))
    out.append(("s_sigma_4", y))
    return out

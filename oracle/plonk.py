"""CPU restatement of the PLONK prover rounds and the verifier (oracle; TEST INFRASTRUCTURE).

PARITY UNPINNED (see oracle/__init__.py): the reference cannot be built here and holds no
golden proof.  This module follows the in-tree round structure line by line --

* key preprocessing        ``src/key.rs:63-327``
* ``create_proof``         ``src/prover.rs:67-474``
* permutation accumulator  ``src/permutation.rs:205-300``
* quotient                 ``src/prover/quotient_poly.rs:20-272``
* linearisation            ``src/prover/linearization_poly.rs:22-225``
* verifier                 ``src/prover/proof.rs:70-591``, ``src/commitment_scheme.rs:24-153``,
                           ``src/verifier.rs:46-81``

-- and restates the gate-widget formulas of the absent ``zksnarks`` crate from upstream
dusk-plonk 0.13 ([EXT-RECALL]).  The widget formulas are corroborated in
three ways by tests/test_oracle_plonk.py: (1) they vanish on every row of circuits built by
the in-tree gadget layouts (``src/lib.rs``), (2) the quotient is a polynomial (exact
division by Z_H) only if they do, (3) the restated verifier -- whose quotient identity
``src/prover/proof.rs:386-440`` *is* in-tree -- accepts the proofs.

Everything is Python big-int arithmetic on canonical values; sizes up to n = 2^9 take
seconds.  Transcript, blinders and the SRS are explicit inputs (they derive from absent
crates in the reference).
"""
from . import curve
from .fields import R_MOD, K1, K2, K3, fr_inv
from .ntt import Fft, poly_eval

_r = R_MOD
EDWARDS_D = (-(10240 * pow(10241, -1, _r))) % _r

SELECTORS = ("q_m", "q_l", "q_r", "q_o", "q_c", "q_d", "q_arith", "q_range", "q_logic",
             "q_fixed_group_add", "q_variable_group_add")

# order in which the verification key seeds the transcript ([EXT-RECALL] dusk-plonk 0.13
# VerifierKey::seed_transcript); selector name -> transcript label
VK_TRANSCRIPT_ORDER = (("q_m", b"q_m"), ("q_l", b"q_l"), ("q_r", b"q_r"), ("q_o", b"q_o"),
                       ("q_c", b"q_c"), ("q_d", b"q_4"), ("q_arith", b"q_arith"),
                       ("q_range", b"q_range"), ("q_logic", b"q_logic"),
                       ("q_variable_group_add", b"q_variable_group_add"),
                       ("q_fixed_group_add", b"q_fixed_group_add"),
                       ("s_sigma_1", b"s_sigma_1"), ("s_sigma_2", b"s_sigma_2"),
                       ("s_sigma_3", b"s_sigma_3"), ("s_sigma_4", b"s_sigma_4"))

EVAL_NAMES = ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
              "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval", "q_arith_eval", "q_c_eval",
              "q_l_eval", "q_r_eval", "perm_eval", "r_poly_eval")


class ProverError(Exception):
    """``Err`` out of ``create_proof``: a commit failed the degree check."""


# ------------------------------------------------------------------ gate widgets
def delta(f):
    return f * (f - 1) % _r * (f - 2) % _r * (f - 3) % _r


def arithmetic_identity(q, a, b, c, d):
    """q = dict of selector values at the row / point."""
    return (a * b % _r * q["q_m"] + a * q["q_l"] + b * q["q_r"] + c * q["q_o"] + d * q["q_d"] + q["q_c"]) % _r \
        * q["q_arith"] % _r


def range_identity(sep, a, b, c, d, d_next):
    k = sep * sep % _r
    k2 = k * k % _r
    k3 = k2 * k % _r
    t = (delta((c - 4 * d) % _r) + delta((b - 4 * c) % _r) * k + delta((a - 4 * b) % _r) * k2
         + delta((d_next - 4 * a) % _r) * k3) % _r
    return t * sep % _r


def _delta_xor_and(a, b, w, c, q_c):
    f = w * (w * (4 * w - 18 * (a + b) + 81) % _r + 18 * (a * a + b * b) - 81 * (a + b) + 83) % _r
    e = (3 * (a + b + c) - 2 * f) % _r
    bb = q_c * (9 * c - 3 * (a + b)) % _r
    return (bb + e) % _r


def logic_identity(sep, a, a_next, b, b_next, c, d, d_next, q_c):
    k = sep * sep % _r
    k2 = k * k % _r
    k3 = k2 * k % _r
    k4 = k3 * k % _r
    A = (a_next - 4 * a) % _r
    B = (b_next - 4 * b) % _r
    D = (d_next - 4 * d) % _r
    c0 = delta(A)
    c1 = delta(B) * k
    c2 = delta(D) * k2
    c3 = (c - A * B) % _r * k3
    c4 = _delta_xor_and(A, B, c, D, q_c) * k4
    return (c3 + c0 + c1 + c2 + c4) % _r * sep % _r


def fixed_base_identity(sep, a, a_next, b, b_next, c, d, d_next, q_l, q_r, q_c):
    k = sep * sep % _r
    k2 = k * k % _r
    k3 = k2 * k % _r
    x_beta, y_beta = q_l, q_r
    acc_x, acc_x_next, acc_y, acc_y_next = a, a_next, b, b_next
    xy_alpha = c
    bit = (d_next - 2 * d) % _r
    bit_consistency = bit * (bit - 1) % _r * (bit + 1) % _r
    y_alpha = (bit * bit % _r * (y_beta - 1) + 1) % _r
    x_alpha = bit * x_beta % _r
    xy_consistency = (bit * q_c - xy_alpha) % _r * k % _r
    t = xy_alpha * acc_x % _r * acc_y % _r * EDWARDS_D % _r
    lhs = (acc_x_next + acc_x_next * t) % _r
    rhs = (acc_x * y_alpha + acc_y * x_alpha) % _r
    x_acc = (lhs - rhs) % _r * k2 % _r
    lhs = (acc_y_next - acc_y_next * t) % _r
    rhs = (acc_y * y_alpha + acc_x * x_alpha) % _r
    y_acc = (lhs - rhs) % _r * k3 % _r
    return (bit_consistency + x_acc + y_acc + xy_consistency) % _r * sep % _r


def var_base_identity(sep, a, a_next, b, b_next, c, d, d_next):
    k = sep * sep % _r
    x_1, x_3, y_1, y_3, x_2, y_2, x1_y2 = a, a_next, b, b_next, c, d, d_next
    xy_consistency = (x_1 * y_2 - x1_y2) % _r
    y1_x2 = y_1 * x_2 % _r
    y1_y2 = y_1 * y_2 % _r
    x1_x2 = x_1 * x_2 % _r
    t = EDWARDS_D * x1_y2 % _r * y1_x2 % _r
    x3c = ((x1_y2 + y1_x2) - (x_3 + x_3 * t)) % _r * k % _r
    y3c = ((y1_y2 + x1_x2) - (y_3 - y_3 * t)) % _r * (k * k % _r) % _r
    return (xy_consistency + x3c + y3c) % _r * sep % _r


def gate_identity(q, seps, a, b, c, d, a_next, b_next, d_next):
    """Sum of the five gate widgets at one row (src/prover/quotient_poly.rs:165-215, without PI)."""
    rs, ls, fs, vs = seps
    t = arithmetic_identity(q, a, b, c, d)
    if q["q_range"]:
        t += q["q_range"] * range_identity(rs, a, b, c, d, d_next)
    if q["q_logic"]:
        t += q["q_logic"] * logic_identity(ls, a, a_next, b, b_next, c, d, d_next, q["q_c"])
    if q["q_fixed_group_add"]:
        t += q["q_fixed_group_add"] * fixed_base_identity(fs, a, a_next, b, b_next, c, d, d_next,
                                                          q["q_l"], q["q_r"], q["q_c"])
    if q["q_variable_group_add"]:
        t += q["q_variable_group_add"] * var_base_identity(vs, a, a_next, b, b_next, c, d, d_next)
    return t % _r


# ------------------------------------------------------------------ circuit plumbing
def wire_values(circ, n):
    """Round-1 gather (src/prover.rs:109-119): four n-vectors, zero beyond m."""
    out = []
    for w in range(4):
        col = [circ.witness[int(i)] for i in circ.wires[w]]
        out.append(col + [0] * (n - len(col)))
    return out


def sigma_evaluations(circ, fft):
    """``compute_permutation_lagrange`` (src/permutation.rs:140-169)."""
    ks = (1, K1, K2, K3)
    roots = fft.elements
    return [[ks[int(w)] * roots[int(g)] % _r for w, g in zip(circ.sigma_w[i], circ.sigma_g[i])]
            for i in range(4)]


class ProvingKey:
    pass


def default_commit(srs_points=None, tau=None):
    """commit(coeffs) -> affine point, raising ProverError on degree overflow.
    With a known tau (synthetic SRS) the MSM collapses to poly(tau) * G."""
    def trim_len(coeffs):
        k = len(coeffs)
        while k and coeffs[k - 1] % _r == 0:
            k -= 1
        return k

    def commit(coeffs, max_len):
        k = trim_len(coeffs)
        if k > max_len:
            raise ProverError("polynomial degree %d exceeds the SRS (%d powers)" % (k - 1, max_len))
        if k == 0:
            return None
        if tau is not None:
            return curve.mul(curve.G1_GEN, poly_eval(coeffs[:k], tau))
        return curve.msm_pippenger(srs_points[:k], coeffs[:k])
    return commit


def compile_circuit(circ, commit, srs_len):
    """``PlonkKey::compile_with_circuit`` (src/key.rs:63-327) -> (ProvingKey, vk dict)."""
    m, n = circ.m, circ.n
    k = n.bit_length() - 1
    fft = Fft(k)
    fft8 = Fft(k + 3)
    pk = ProvingKey()
    pk.n, pk.m, pk.k = n, m, k
    pk.fft, pk.fft8 = fft, fft8
    additional_n = 1 << (m + 6 - 1).bit_length()
    pk.max_len = min(srs_len, additional_n + 7)  # trim(additional_n) keeps room for t_4
    pk.poly, pk.eval8 = {}, {}
    for s in SELECTORS:
        col = [int(v) % _r for v in circ.selectors[s]] + [0] * (n - m)
        pk.poly[s] = fft.idft(col)
    pk.sigma_evals = sigma_evaluations(circ, fft)
    for i in range(4):
        pk.poly["s_sigma_%d" % (i + 1)] = fft.idft(pk.sigma_evals[i])
    vk = {"n": m, "n_inv": fft.size_inv(), "generator": fft.generator(), "generator_inv": fft.generator_inv()}
    for s in SELECTORS:
        try:
            vk[s] = commit(pk.poly[s], pk.max_len)   # .unwrap_or_default()
        except ProverError:
            vk[s] = None
    for i in range(4):
        nm = "s_sigma_%d" % (i + 1)
        vk[nm] = commit(pk.poly[nm], pk.max_len)
    for nm, p in pk.poly.items():
        pk.eval8[nm] = fft8.coset_dft(p)
    pk.eval8["linear"] = fft8.coset_dft([0, 1])
    pk.v_h_coset_8n = fft8.compute_vanishing_poly_over_coset(n)
    pk.vk = vk
    return pk, vk


def blind(poly, blinders, n):
    """``Coefficients::blind(h, rng)`` with the h+1 scalars given: (b0 + b1 X + ..)(X^n - 1)."""
    out = list(poly) + [0] * (n - len(poly))
    for i, b in enumerate(blinders):
        out[i] = (out[i] - b) % _r
        out.append(b % _r)
    return out


def compute_permutation_vec(fft, wires, beta, gamma, sigma_evals):
    """src/permutation.rs:205-300."""
    n = fft.size()
    ks = (1, K1, K2, K3)
    roots = fft.elements
    z = [1]
    state = 1
    for i in range(n):
        num = den = 1
        for j in range(4):
            num = num * ((wires[j][i] + beta * ks[j] % _r * roots[i] + gamma) % _r) % _r
            den = den * ((wires[j][i] + beta * sigma_evals[j][i] + gamma) % _r) % _r
        state = state * num % _r * fr_inv(den) % _r
        z.append(state)
    z.pop()
    return z


def compute_quotient(pk, z_poly, w_polys, pi_poly, ch):
    """src/prover/quotient_poly.rs:20-118 -> 8n coefficients."""
    alpha, beta, gamma, rs, ls, fs, vs = ch
    fft, fft8 = pk.fft, pk.fft8
    n8 = fft8.size()
    ev = lambda p: fft8.coset_dft(p)
    z8, a8, b8, c8, d8 = ev(z_poly), ev(w_polys[0]), ev(w_polys[1]), ev(w_polys[2]), ev(w_polys[3])
    for v in (z8, a8, b8, d8):
        v.extend(v[:8])
    pi8 = ev(pi_poly)
    e = pk.eval8
    # t_1: gate identities + PI (src/prover/quotient_poly.rs:154-217)
    t1 = []
    for i in range(n8):
        q = {s: e[s][i] for s in SELECTORS}
        g = gate_identity(q, (rs, ls, fs, vs), a8[i], b8[i], c8[i], d8[i], a8[i + 8], b8[i + 8], d8[i + 8])
        t1.append((g + pi8[i]) % _r)
    # t_2: permutation (src/prover/quotient_poly.rs:222-262)
    l1 = fft.idft([alpha * alpha % _r] + [0] * (fft.size() - 1))
    l18 = ev(l1)
    t2 = []
    for i in range(n8):
        x = e["linear"][i]
        ident = (a8[i] + beta * x + gamma) % _r * ((b8[i] + beta * K1 % _r * x + gamma) % _r) % _r \
            * ((c8[i] + beta * K2 % _r * x + gamma) % _r) % _r * ((d8[i] + beta * K3 % _r * x + gamma) % _r) % _r \
            * z8[i] % _r * alpha % _r
        copy = (a8[i] + beta * e["s_sigma_1"][i] + gamma) % _r * ((b8[i] + beta * e["s_sigma_2"][i] + gamma) % _r) % _r \
            * ((c8[i] + beta * e["s_sigma_3"][i] + gamma) % _r) % _r \
            * ((d8[i] + beta * e["s_sigma_4"][i] + gamma) % _r) % _r * z8[i + 8] % _r * alpha % _r
        one = (z8[i] - 1) * l18[i] % _r
        t2.append((ident - copy + one) % _r)
    quotient = [(t1[i] + t2[i]) * fr_inv(pk.v_h_coset_8n[i]) % _r for i in range(n8)]
    return fft8.coset_idft(quotient)


def padd(a, b):
    if len(a) < len(b):
        a, b = b, a
    return [(x + (b[i] if i < len(b) else 0)) % _r for i, x in enumerate(a)]


def pscale(a, s):
    return [x * s % _r for x in a]


def compute_linearization(pk, ch, w_polys, t_poly, z_poly):
    """src/prover/linearization_poly.rs:22-134 -> (r_poly, evaluations dict, t_eval)."""
    alpha, beta, gamma, rs, ls, fs, vs, zc = ch
    P = pk.poly
    gen = pk.fft.generator()
    zw = zc * gen % _r
    ev = {}
    t_eval = poly_eval(t_poly, zc)
    a, b, c, d = (poly_eval(p, zc) for p in w_polys)
    ev["a_eval"], ev["b_eval"], ev["c_eval"], ev["d_eval"] = a, b, c, d
    s1, s2, s3 = (poly_eval(P["s_sigma_%d" % i], zc) for i in (1, 2, 3))
    ev["s_sigma_1_eval"], ev["s_sigma_2_eval"], ev["s_sigma_3_eval"] = s1, s2, s3
    ev["q_arith_eval"] = qa = poly_eval(P["q_arith"], zc)
    ev["q_c_eval"] = qc = poly_eval(P["q_c"], zc)
    ev["q_l_eval"] = ql = poly_eval(P["q_l"], zc)
    ev["q_r_eval"] = qr = poly_eval(P["q_r"], zc)
    ev["a_next_eval"] = an = poly_eval(w_polys[0], zw)
    ev["b_next_eval"] = bn = poly_eval(w_polys[1], zw)
    ev["d_next_eval"] = dn = poly_eval(w_polys[3], zw)
    ev["perm_eval"] = pe = poly_eval(z_poly, zw)
    scal = linearization_scalars(pk.n, ch, ev)
    r = []
    for name, s in scal:
        src = z_poly if name == "z" else P[name]
        r = padd(r, pscale(src, s)) if r else pscale(src, s)
    ev["r_poly_eval"] = poly_eval(r, zc)
    return r, ev, t_eval


def linearization_scalars(n, ch, ev):
    """(polynomial name, scalar) pairs with r(X) = sum scalar * poly(X); the same pairs give
    the verifier's linearisation commitment (src/prover/proof.rs:459-527)."""
    alpha, beta, gamma, rs, ls, fs, vs, zc = ch
    a, b, c, d = ev["a_eval"], ev["b_eval"], ev["c_eval"], ev["d_eval"]
    an, bn, dn = ev["a_next_eval"], ev["b_next_eval"], ev["d_next_eval"]
    qa, qc, ql, qr = ev["q_arith_eval"], ev["q_c_eval"], ev["q_l_eval"], ev["q_r_eval"]
    s1, s2, s3, pe = ev["s_sigma_1_eval"], ev["s_sigma_2_eval"], ev["s_sigma_3_eval"], ev["perm_eval"]
    out = [("q_m", a * b % _r * qa % _r), ("q_l", a * qa % _r), ("q_r", b * qa % _r), ("q_o", c * qa % _r),
           ("q_d", d * qa % _r), ("q_c", qa),
           ("q_range", range_identity(rs, a, b, c, d, dn)),
           ("q_logic", logic_identity(ls, a, an, b, bn, c, d, dn, qc)),
           ("q_fixed_group_add", fixed_base_identity(fs, a, an, b, bn, c, d, dn, ql, qr, qc)),
           ("q_variable_group_add", var_base_identity(vs, a, an, b, bn, c, d, dn))]
    # permutation.linearize
    zh = (pow(zc, n, _r) - 1) % _r
    l1 = zh * fr_inv(n * (zc - 1) % _r) % _r
    x = (a + beta * zc + gamma) % _r * ((b + beta * K1 % _r * zc + gamma) % _r) % _r \
        * ((c + beta * K2 % _r * zc + gamma) % _r) % _r * ((d + beta * K3 % _r * zc + gamma) % _r) % _r * alpha % _r
    y = (-((a + beta * s1 + gamma) % _r * ((b + beta * s2 + gamma) % _r) % _r * ((c + beta * s3 + gamma) % _r) % _r
           * beta % _r * pe % _r * alpha)) % _r
    out.append(("z", (x + l1 * alpha % _r * alpha) % _r))
    out.append(("s_sigma_4", y))
    return out


def ruffini(poly, point):
    """Synthetic division by (X - point), remainder dropped."""
    q = [0] * (len(poly) - 1) if poly else []
    carry = 0
    for i in range(len(poly) - 1, 0, -1):
        carry = (poly[i] + carry * point) % _r
        q[i - 1] = carry
    return q


def compute_aggregate_witness(polys, point, v):
    num = []
    pw = 1
    for p in polys:
        num = padd(num, pscale(p, pw)) if num else pscale(p, pw)
        pw = pw * v % _r
    return ruffini(num, point)


class Proof:
    COMM_NAMES = ("a_comm", "b_comm", "c_comm", "d_comm", "z_comm", "t_low_comm", "t_mid_comm",
                  "t_high_comm", "t_4_comm", "w_z_chall_comm", "w_z_chall_w_comm")

    def __init__(self):
        self.evaluations = {}

    def __eq__(self, o):
        return all(getattr(self, c) == getattr(o, c) for c in self.COMM_NAMES) and self.evaluations == o.evaluations

    # field order of ``Evaluations`` (src/prover/linearization_poly.rs:113-130)
    WIRE_EVAL_ORDER = ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
                       "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "s_sigma_1_eval", "s_sigma_2_eval",
                       "s_sigma_3_eval", "r_poly_eval", "perm_eval")

    def to_bytes(self):
        """1040-byte wire form (src/prover/proof.rs:36-66 derives SCALE Encode: fixed-size fields in declaration
        order): 11 compressed G1 then 16 little-endian scalars.  Oracle-side serialiser (own code path)."""
        from .merlin import compress_g1
        out = b"".join(compress_g1(getattr(self, c)) for c in self.COMM_NAMES)
        return out + b"".join((self.evaluations[k] % R_MOD).to_bytes(32, "little") for k in self.WIRE_EVAL_ORDER)


def create_proof(pk, circ, commit, transcript, blinders, trace=None):
    """``Prover::create_proof`` (src/prover.rs:67-474).  ``blinders``: 4 x 2 + 3 scalars in
    the RNG draw order a, b, o, d, z.  ``transcript`` is the prover's base transcript (cloned).
    ``trace``: optional dict receiving every intermediate polynomial for parity tests."""
    n = pk.n
    fft = pk.fft
    tr = transcript.clone()
    T = trace if trace is not None else {}
    cm = lambda p: commit(p, pk.max_len)
    for pi in circ.pi_values:
        tr.append_scalar(b"pi", pi)
    # round 1
    w_scalar = wire_values(circ, n)
    w_polys = [blind(fft.idft(w_scalar[i]), blinders[2 * i:2 * i + 2], n) for i in range(4)]
    T["w_polys"] = w_polys
    proof = Proof()
    proof.a_comm, proof.b_comm, proof.c_comm, proof.d_comm = (cm(p) for p in w_polys)
    for lab, c in ((b"a_w", proof.a_comm), (b"b_w", proof.b_comm), (b"c_w", proof.c_comm), (b"d_w", proof.d_comm)):
        tr.append_commitment(lab, c)
    # round 2
    beta = tr.challenge_scalar(b"beta")
    tr.append_scalar(b"beta", beta)
    gamma = tr.challenge_scalar(b"gamma")
    sigma_evals = [fft.dft(pk.poly["s_sigma_%d" % (i + 1)]) for i in range(4)]
    zvec = compute_permutation_vec(fft, w_scalar, beta, gamma, sigma_evals)
    T["z_evals"] = zvec
    z_poly = blind(fft.idft(zvec), blinders[8:11], n)
    T["z_poly"] = z_poly
    proof.z_comm = cm(z_poly)
    tr.append_commitment(b"z", proof.z_comm)
    # round 3
    alpha = tr.challenge_scalar(b"alpha")
    rs = tr.challenge_scalar(b"range separation challenge")
    ls = tr.challenge_scalar(b"logic separation challenge")
    fs = tr.challenge_scalar(b"fixed base separation challenge")
    vs = tr.challenge_scalar(b"variable base separation challenge")
    dense_pi = [0] * n
    for i, v in zip(circ.pi_indexes, circ.pi_values):
        dense_pi[i] = v
    pi_poly = fft.idft(dense_pi)
    ch7 = (alpha, beta, gamma, rs, ls, fs, vs)
    T["challenges"] = ch7
    t_poly = compute_quotient(pk, z_poly, w_polys, pi_poly, ch7)
    T["t_poly"] = t_poly
    t_low, t_mid, t_high, t_4 = t_poly[0:n], t_poly[n:2 * n], t_poly[2 * n:3 * n], t_poly[3 * n:]
    proof.t_low_comm, proof.t_mid_comm, proof.t_high_comm, proof.t_4_comm = cm(t_low), cm(t_mid), cm(t_high), cm(t_4)
    for lab, c in ((b"t_low", proof.t_low_comm), (b"t_mid", proof.t_mid_comm), (b"t_high", proof.t_high_comm),
                   (b"t_4", proof.t_4_comm)):
        tr.append_commitment(lab, c)
    # round 4 / 5
    zc = tr.challenge_scalar(b"z_challenge")
    T["z_challenge"] = zc
    r_poly, ev, t_eval = compute_linearization(pk, ch7 + (zc,), w_polys, t_poly, z_poly)
    T["r_poly"] = r_poly
    T["t_eval"] = t_eval
    for nm in ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
               "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval", "q_arith_eval", "q_c_eval", "q_l_eval",
               "q_r_eval", "perm_eval"):
        tr.append_scalar(nm.encode(), ev[nm])
    tr.append_scalar(b"t_eval", t_eval)
    tr.append_scalar(b"r_eval", ev["r_poly_eval"])
    z_n = pow(zc, n, _r)
    quot = padd(padd(padd(t_low, pscale(t_mid, z_n)), pscale(t_high, z_n * z_n % _r)),
                pscale(t_4, pow(z_n, 3, _r)))
    v1 = tr.challenge_scalar(b"v_challenge")
    agg = compute_aggregate_witness(
        [quot, r_poly, w_polys[0], w_polys[1], w_polys[2], w_polys[3],
         pk.poly["s_sigma_1"], pk.poly["s_sigma_2"], pk.poly["s_sigma_3"]], zc, v1)
    T["w_z_poly"] = agg
    proof.w_z_chall_comm = cm(agg)
    v2 = tr.challenge_scalar(b"v_challenge")
    sagg = compute_aggregate_witness([z_poly, w_polys[0], w_polys[1], w_polys[3]],
                                     zc * fft.generator() % _r, v2)
    T["w_zw_poly"] = sagg
    proof.w_z_chall_w_comm = cm(sagg)
    proof.evaluations = {k: ev[k] for k in EVAL_NAMES}
    return proof, list(circ.pi_values)


# ------------------------------------------------------------------ verifier
class VerifyError(Exception):
    pass


def _g1_lincomb(pairs):
    acc = curve.J_INF
    for s, pt in pairs:
        if pt is None or s % _r == 0:
            continue
        acc = curve.jadd(acc, curve.jmul(curve.to_jac(pt), s))
    return curve.to_affine(acc)


def verify(vk, n, proof, pi_indexes, pi_values, transcript, kzg_check):
    """``Verifier::verify`` + ``Proof::verify`` (src/verifier.rs:46-81, src/prover/proof.rs:70-383).
    ``kzg_check(total_w_neg, total_c) -> bool`` performs the final 2-pairing product check
    of ``batch_check`` (src/commitment_scheme.rs:52-64)."""
    if len(pi_values) != len(pi_indexes):
        raise VerifyError("InconsistentPublicInputsLen")
    tr = transcript.clone()
    for pi in pi_values:
        tr.append_scalar(b"pi", pi)
    dense = [0] * n
    for i, v in zip(pi_indexes, pi_values):
        dense[i] = v
    E = proof.evaluations
    for lab, c in ((b"a_w", proof.a_comm), (b"b_w", proof.b_comm), (b"c_w", proof.c_comm), (b"d_w", proof.d_comm)):
        tr.append_commitment(lab, c)
    beta = tr.challenge_scalar(b"beta")
    tr.append_scalar(b"beta", beta)
    gamma = tr.challenge_scalar(b"gamma")
    tr.append_commitment(b"z", proof.z_comm)
    alpha = tr.challenge_scalar(b"alpha")
    rs = tr.challenge_scalar(b"range separation challenge")
    ls = tr.challenge_scalar(b"logic separation challenge")
    fs = tr.challenge_scalar(b"fixed base separation challenge")
    vs = tr.challenge_scalar(b"variable base separation challenge")
    for lab, c in ((b"t_low", proof.t_low_comm), (b"t_mid", proof.t_mid_comm), (b"t_high", proof.t_high_comm),
                   (b"t_4", proof.t_4_comm)):
        tr.append_commitment(lab, c)
    zc = tr.challenge_scalar(b"z_challenge")
    assert n == 1 << (vk["n"] - 1).bit_length()
    n_inv, gen, gen_inv = vk["n_inv"], vk["generator"], vk["generator_inv"]
    z_h = (pow(zc, n, _r) - 1) % _r
    l1 = z_h * fr_inv(n * (zc - 1) % _r) % _r
    # barycentric PI(z)  (src/prover/proof.rs:541-591)
    num = z_h * n_inv % _r
    pi_eval = 0
    for i, v in enumerate(dense):
        if v:
            pi_eval = (pi_eval + v * fr_inv((pow(gen_inv, i, _r) * zc - 1) % _r)) % _r
    pi_eval = pi_eval * num % _r
    a = (E["r_poly_eval"] + pi_eval) % _r
    b = (E["a_eval"] + beta * E["s_sigma_1_eval"] + gamma) % _r * ((E["b_eval"] + beta * E["s_sigma_2_eval"] + gamma) % _r) % _r \
        * ((E["c_eval"] + beta * E["s_sigma_3_eval"] + gamma) % _r) % _r \
        * ((E["d_eval"] + gamma) % _r * E["perm_eval"] % _r * alpha % _r) % _r
    c = l1 * alpha % _r * alpha % _r
    t_eval = (a - b - c) % _r * fr_inv(z_h) % _r
    z_n = pow(zc, n, _r)
    t_comm = _g1_lincomb([(1, proof.t_low_comm), (z_n, proof.t_mid_comm), (z_n * z_n % _r, proof.t_high_comm),
                          (pow(z_n, 3, _r), proof.t_4_comm)])
    for nm in ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
               "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval", "q_arith_eval", "q_c_eval", "q_l_eval",
               "q_r_eval", "perm_eval"):
        tr.append_scalar(nm.encode(), E[nm])
    tr.append_scalar(b"t_eval", t_eval)
    tr.append_scalar(b"r_eval", E["r_poly_eval"])
    scal = linearization_scalars(n, (alpha, beta, gamma, rs, ls, fs, vs, zc), E)
    r_comm = _g1_lincomb([(s, proof.z_comm if nm == "z" else vk[nm]) for nm, s in scal])

    def flatten(witness, parts):
        v = tr.challenge_scalar(b"v_challenge")
        pw, pts, evs = 1, [], 0
        for e, cmt in parts:
            pts.append((pw, cmt))
            evs = (evs + e * pw) % _r
            pw = pw * v % _r
        return witness, evs, _g1_lincomb(pts)

    pa = flatten(proof.w_z_chall_comm, [
        (t_eval, t_comm), (E["r_poly_eval"], r_comm), (E["a_eval"], proof.a_comm), (E["b_eval"], proof.b_comm),
        (E["c_eval"], proof.c_comm), (E["d_eval"], proof.d_comm), (E["s_sigma_1_eval"], vk["s_sigma_1"]),
        (E["s_sigma_2_eval"], vk["s_sigma_2"]), (E["s_sigma_3_eval"], vk["s_sigma_3"])])
    pb = flatten(proof.w_z_chall_w_comm, [
        (E["perm_eval"], proof.z_comm), (E["a_next_eval"], proof.a_comm), (E["b_next_eval"], proof.b_comm),
        (E["d_next_eval"], proof.d_comm)])
    tr.append_commitment(b"w_z", proof.w_z_chall_comm)
    tr.append_commitment(b"w_z_w", proof.w_z_chall_w_comm)
    # batch_check (src/commitment_scheme.rs:24-66)
    u = tr.challenge_scalar(b"batch")
    total_c, total_w = [], []
    g_mult = 0
    pw = 1
    for (w, ev, cm_), point in zip((pa, pb), (zc, zc * gen % _r)):
        total_c.append((pw, cm_))
        total_c.append((pw * point % _r, w))
        g_mult = (g_mult + pw * ev) % _r
        total_w.append((pw, w))
        pw = pw * u % _r
    total_c.append(((-g_mult) % _r, curve.G1_GEN))
    tc = _g1_lincomb(total_c)
    tw = curve.neg(_g1_lincomb(total_w))
    if not kzg_check(tw, tc):
        raise VerifyError("ProofVerificationError")
    return True


def trapdoor_kzg_check(tau):
    """e(W, [tau]_2) * e(C, [1]_2) == 1  <=>  tau * W + C == O when tau is known (synthetic SRS)."""
    def check(total_w_neg, total_c):
        return curve.add(curve.mul(total_w_neg, tau), total_c) is None
    return check


def vk_transcript_list(vk):
    return [(lab, vk[name]) for name, lab in VK_TRANSCRIPT_ORDER]

"""Full-size CPU prover on the C restatement (oracle; TEST INFRASTRUCTURE and CPU baseline).

Same round structure as oracle/plonk.py (which follows src/key.rs:63-327 and
src/prover.rs:67-474 line by line) but on numpy limb arrays with the hot loops in
oracle/zkp_oracle.c / zkp_oracle_prover.c (OpenMP), so 2^16-gate circuits prove in seconds.
It is validated against the pure-Python prover on small circuits (tests/test_oracle_c.py), is
the full-size parity oracle for the GPU prover and is what ``bench.py`` times as the
``"port"`` CPU baseline.  PARITY UNPINNED: see oracle/__init__.py.
"""
import ctypes

import numpy as np

from . import cport
from .fields import (R_MOD, K1, K2, K3, fr_from_mont_limbs, fr_to_mont_limbs, g1_from_mont_limbs)
from .plonk import (SELECTORS, EVAL_NAMES, VK_TRANSCRIPT_ORDER, Proof, ProverError, linearization_scalars)

_r = R_MOD
_u64p = ctypes.POINTER(ctypes.c_uint64)
_sz = ctypes.c_size_t


def _p(a):
    return a.ctypes.data_as(_u64p)


def _lib():
    lib = cport.lib()
    if not getattr(lib, "_prover_bound", False):
        lib.oracle_perm_z.argtypes = [_u64p, _u64p, _u64p, _u64p, _u64p, _sz, _u64p, ctypes.c_int, ctypes.c_int]
        lib.oracle_quotient.argtypes = [ctypes.POINTER(_u64p), _u64p, _sz, ctypes.c_int, _u64p, ctypes.c_int]
        lib.oracle_poly_eval.argtypes = [_u64p, _sz, _u64p, _u64p, ctypes.c_int]
        lib.oracle_poly_lincomb.argtypes = [ctypes.POINTER(_u64p), _u64p, _u64p, ctypes.c_uint, _u64p, _sz, ctypes.c_int]
        lib.oracle_poly_div_linear.argtypes = [_u64p, _sz, _u64p, _u64p]
        lib._prover_bound = True
    return lib


def m1(v):
    return fr_to_mont_limbs([v])[0]


def column_to_mont(col, n):
    out = np.zeros((n, 4), dtype=np.uint64)
    if isinstance(col, np.ndarray) and col.dtype != object:
        uniq, inv = np.unique(col, return_inverse=True)
        out[:len(col)] = fr_to_mont_limbs([int(u) for u in uniq])[inv]
        return out
    cache, keys = {}, []
    idx = np.empty(len(col), dtype=np.int64)
    for i, v in enumerate(col):
        j = cache.get(v)
        if j is None:
            j = cache[v] = len(keys)
            keys.append(v)
        idx[i] = j
    if keys:
        out[:len(col)] = fr_to_mont_limbs(keys)[idx]
    return out


def perm_z(wires, sigmas, roots, beta, gamma, faithful=False):
    n = roots.shape[0]
    out = np.zeros((n, 4), dtype=np.uint64)
    w = np.ascontiguousarray(wires).reshape(4 * n, 4)
    s = np.ascontiguousarray(sigmas).reshape(4 * n, 4)
    rc = _lib().oracle_perm_z(_p(w), _p(s), _p(roots), _p(m1(beta)), _p(m1(gamma)), n, _p(out), int(faithful), 0)
    assert rc == 0
    return out


def quotient(cols, challenges, n8, faithful=False):
    assert len(cols) == 24
    cols = [np.ascontiguousarray(c) for c in cols]
    arr = (_u64p * 24)(*[_p(c) for c in cols])
    ch = fr_to_mont_limbs(challenges)
    out = np.zeros((n8, 4), dtype=np.uint64)
    rc = _lib().oracle_quotient(arr, _p(ch), n8, int(faithful), _p(out), 0)
    assert rc == 0
    return out


def poly_eval(p, point):
    p = np.ascontiguousarray(p)
    out = np.zeros(4, dtype=np.uint64)
    _lib().oracle_poly_eval(_p(p), p.shape[0], _p(m1(point)), _p(out), 0)
    return fr_from_mont_limbs(out)[0]


def poly_lincomb(polys, scalars, out_len):
    polys = [np.ascontiguousarray(p) for p in polys]
    arr = (_u64p * len(polys))(*[_p(p) for p in polys])
    lens = np.array([p.shape[0] for p in polys], dtype=np.uint64)
    sc = fr_to_mont_limbs(scalars)
    out = np.zeros((out_len, 4), dtype=np.uint64)
    _lib().oracle_poly_lincomb(arr, _p(lens), _p(sc), len(polys), _p(out), out_len, 0)
    return out


def poly_div_linear(p, point):
    p = np.ascontiguousarray(p)
    out = np.zeros((max(p.shape[0] - 1, 0), 4), dtype=np.uint64)
    if p.shape[0] > 1:
        _lib().oracle_poly_div_linear(_p(p), p.shape[0], _p(m1(point)), _p(out))
    return out


def blind(poly, blinders, n):
    out = np.zeros((n + len(blinders), 4), dtype=np.uint64)
    out[:n] = poly[:n]
    head = fr_from_mont_limbs(out[:len(blinders)])
    out[:len(blinders)] = fr_to_mont_limbs([(h - b) % _r for h, b in zip(head, blinders)])
    out[n:] = fr_to_mont_limbs(blinders)
    return out


class CProver:
    """compile + create_proof on the CPU.  ``srs_xy``: (N, 12) uint64 affine powers."""

    def __init__(self, circ, srs_xy, label, transcript_cls, faithful=False):
        self.circ_m, self.n = circ.m, circ.n
        n = self.n
        self.k = k = n.bit_length() - 1
        self.faithful = faithful
        additional_n = 1 << (circ.m + 6 - 1).bit_length()
        self.max_len = min(srs_xy.shape[0], additional_n + 7)
        self.srs = np.ascontiguousarray(srs_xy[:self.max_len])
        self.poly, self.eval8 = {}, {}
        for s in SELECTORS:
            self.poly[s] = cport.ntt(column_to_mont(circ.selectors[s], n), k, inverse=True)
        # roots and sigma evaluations
        w = pow(7, (R_MOD - 1) >> k, R_MOD) if k else 1
        roots, x = [], 1
        for _ in range(n):
            roots.append(x)
            x = x * w % _r
        self.roots = fr_to_mont_limbs(roots)
        ks = (1, K1, K2, K3)
        kroots = [fr_to_mont_limbs([kk * v % _r for v in roots]) for kk in ks]
        self.sigma_evals = []
        for i in range(4):
            sw = np.asarray(circ.sigma_w[i], dtype=np.int64)
            sg = np.asarray(circ.sigma_g[i], dtype=np.int64)
            ev = np.zeros((n, 4), dtype=np.uint64)
            for kk in range(4):
                msk = sw == kk
                ev[msk] = kroots[kk][sg[msk]]
            self.sigma_evals.append(ev)
            self.poly["s_sigma_%d" % (i + 1)] = cport.ntt(ev, k, inverse=True)
        self.vk = {"n": circ.m, "n_inv": pow(n, -1, _r), "generator": w, "generator_inv": pow(w, -1, _r)}
        for s in SELECTORS:
            try:
                self.vk[s] = self.commit(self.poly[s])
            except ProverError:
                self.vk[s] = None
        for i in range(4):
            nm = "s_sigma_%d" % (i + 1)
            self.vk[nm] = self.commit(self.poly[nm])
        for nm, p in self.poly.items():
            self.eval8[nm] = cport.ntt(p, k + 3, coset=True)
        self.eval8["linear"] = cport.ntt(fr_to_mont_limbs([0, 1]), k + 3, coset=True)
        w8 = pow(7, (R_MOD - 1) >> (k + 3), R_MOD)
        gn, wn = pow(7, n, _r), pow(w8, n, _r)
        vh8 = fr_to_mont_limbs([(gn * pow(wn, i, _r) - 1) % _r for i in range(8)])
        self.v_h_coset_8n = np.ascontiguousarray(np.tile(vh8, (n, 1)))
        self.transcript = transcript_cls.base(label, [(lab, self.vk[name]) for name, lab in VK_TRANSCRIPT_ORDER],
                                              circ.m)

    def commit(self, coeffs):
        nz = np.nonzero(coeffs.any(axis=1))[0]
        top = int(nz[-1]) + 1 if len(nz) else 0
        if top > self.max_len:
            raise ProverError("polynomial degree %d exceeds the SRS (%d powers)" % (top - 1, self.max_len))
        if top == 0:
            return None
        return g1_from_mont_limbs(cport.msm_g1(self.srs, coeffs[:top]))[0]

    def create_proof(self, blinders, circ, trace=None):
        n, k = self.n, self.k
        T = trace if trace is not None else {}
        tr = self.transcript.clone()
        for pi in circ.pi_values:
            tr.append_scalar(b"pi", pi)
        wit = fr_to_mont_limbs(circ.witness)
        idx = np.asarray(circ.wires, dtype=np.int64)
        W = np.zeros((4, n, 4), dtype=np.uint64)
        for j in range(4):
            W[j, :idx.shape[1]] = wit[idx[j]]
        wp = [blind(cport.ntt(W[j], k, inverse=True), blinders[2 * j:2 * j + 2], n) for j in range(4)]
        T["w_polys"] = wp
        proof = Proof()
        proof.a_comm, proof.b_comm, proof.c_comm, proof.d_comm = (self.commit(p) for p in wp)
        for lab, c in ((b"a_w", proof.a_comm), (b"b_w", proof.b_comm), (b"c_w", proof.c_comm), (b"d_w", proof.d_comm)):
            tr.append_commitment(lab, c)
        beta = tr.challenge_scalar(b"beta")
        tr.append_scalar(b"beta", beta)
        gamma = tr.challenge_scalar(b"gamma")
        sig = [cport.ntt(self.poly["s_sigma_%d" % (i + 1)], k) for i in range(4)]   # permutation.rs:229-235
        zv = perm_z(W, np.stack(sig), self.roots, beta, gamma, self.faithful)
        T["z_evals"] = zv
        zp = blind(cport.ntt(zv, k, inverse=True), blinders[8:11], n)
        T["z_poly"] = zp
        proof.z_comm = self.commit(zp)
        tr.append_commitment(b"z", proof.z_comm)
        alpha = tr.challenge_scalar(b"alpha")
        rs = tr.challenge_scalar(b"range separation challenge")
        ls = tr.challenge_scalar(b"logic separation challenge")
        fs = tr.challenge_scalar(b"fixed base separation challenge")
        vs = tr.challenge_scalar(b"variable base separation challenge")
        ch7 = (alpha, beta, gamma, rs, ls, fs, vs)
        T["challenges"] = ch7
        dense = np.zeros((n, 4), dtype=np.uint64)
        if len(circ.pi_indexes):
            dense[np.asarray(circ.pi_indexes, dtype=np.int64)] = fr_to_mont_limbs(circ.pi_values)
        pi_poly = cport.ntt(dense, k, inverse=True)
        k8, n8 = k + 3, 8 * n
        e8 = [cport.ntt(p, k8, coset=True) for p in wp]
        z8 = cport.ntt(zp, k8, coset=True)
        pi8 = cport.ntt(pi_poly, k8, coset=True)
        l1 = np.zeros((n, 4), dtype=np.uint64)
        l1[0] = m1(alpha * alpha % _r)
        l18 = cport.ntt(cport.ntt(l1, k, inverse=True), k8, coset=True)
        cols = e8 + [z8, pi8, l18] + [self.eval8[s] for s in SELECTORS] + \
            [self.eval8["s_sigma_%d" % i] for i in (1, 2, 3, 4)] + [self.eval8["linear"], self.v_h_coset_8n]
        t8 = quotient(cols, ch7, n8, self.faithful)
        t_poly = cport.ntt(t8, k8, inverse=True, coset=True)
        T["t_poly"] = t_poly
        parts = [t_poly[0:n], t_poly[n:2 * n], t_poly[2 * n:3 * n], t_poly[3 * n:]]
        tc = [self.commit(p) for p in parts]
        proof.t_low_comm, proof.t_mid_comm, proof.t_high_comm, proof.t_4_comm = tc
        for lab, c in zip((b"t_low", b"t_mid", b"t_high", b"t_4"), tc):
            tr.append_commitment(lab, c)
        zc = tr.challenge_scalar(b"z_challenge")
        T["z_challenge"] = zc
        zw = zc * self.vk["generator"] % _r
        P = self.poly
        ev = {}
        t_eval = poly_eval(t_poly, zc)
        ev["a_eval"], ev["b_eval"], ev["c_eval"], ev["d_eval"] = (poly_eval(p, zc) for p in wp)
        for i in (1, 2, 3):
            ev["s_sigma_%d_eval" % i] = poly_eval(P["s_sigma_%d" % i], zc)
        ev["q_arith_eval"] = poly_eval(P["q_arith"], zc)
        ev["q_c_eval"] = poly_eval(P["q_c"], zc)
        ev["q_l_eval"] = poly_eval(P["q_l"], zc)
        ev["q_r_eval"] = poly_eval(P["q_r"], zc)
        ev["a_next_eval"] = poly_eval(wp[0], zw)
        ev["b_next_eval"] = poly_eval(wp[1], zw)
        ev["d_next_eval"] = poly_eval(wp[3], zw)
        ev["perm_eval"] = poly_eval(zp, zw)
        scal = linearization_scalars(n, ch7 + (zc,), ev)
        r_poly = poly_lincomb([zp if nm == "z" else P[nm] for nm, _ in scal], [s for _, s in scal], n + 3)
        T["r_poly"] = r_poly
        T["t_eval"] = t_eval
        ev["r_poly_eval"] = poly_eval(r_poly, zc)
        for nm in ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
                   "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval", "q_arith_eval", "q_c_eval", "q_l_eval",
                   "q_r_eval", "perm_eval"):
            tr.append_scalar(nm.encode(), ev[nm])
        tr.append_scalar(b"t_eval", t_eval)
        tr.append_scalar(b"r_eval", ev["r_poly_eval"])
        z_n = pow(zc, n, _r)
        v1 = tr.challenge_scalar(b"v_challenge")
        quot = poly_lincomb(parts, [1, z_n, z_n * z_n % _r, pow(z_n, 3, _r)], 5 * n)
        agg = poly_lincomb([quot, r_poly, wp[0], wp[1], wp[2], wp[3], P["s_sigma_1"], P["s_sigma_2"], P["s_sigma_3"]],
                           [pow(v1, i, _r) for i in range(9)], 5 * n)
        wz = poly_div_linear(agg, zc)
        T["w_z_poly"] = wz
        proof.w_z_chall_comm = self.commit(wz)
        v2 = tr.challenge_scalar(b"v_challenge")
        sagg = poly_lincomb([zp, wp[0], wp[1], wp[3]], [pow(v2, i, _r) for i in range(4)], n + 3)
        wzw = poly_div_linear(sagg, zw)
        T["w_zw_poly"] = wzw
        proof.w_z_chall_w_comm = self.commit(wzw)
        proof.evaluations = {nm: ev[nm] for nm in EVAL_NAMES}
        return proof, list(circ.pi_values)

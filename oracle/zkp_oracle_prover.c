/*
 * CPU restatement of the element-wise prover rounds -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * PARITY UNPINNED (see zkp_oracle.c): the reference cannot be built and holds no golden proof.
 *
 * Follows, loop for loop:
 *   oracle_perm_z          Permutation::compute_permutation_vec     src/permutation.rs:205-300
 *   oracle_quotient        quotient_poly::compute between its NTTs  src/prover/quotient_poly.rs:74-114,
 *                          compute_circuit_satisfiability_equation  :122-219, compute_permutation_checks :222-262
 *   oracle_poly_eval       Coefficients::evaluate                   src/prover/linearization_poly.rs:52-73
 *   oracle_poly_lincomb    &poly * &scalar + ..                     src/prover.rs:408-418, linearization_poly.rs:75-105
 *   oracle_poly_div_linear ruffini in compute_aggregate_witness     src/prover.rs:422-451
 * Gate-widget formulas: [EXT-RECALL] dusk-plonk 0.13; checked against the
 * Python restatement (oracle/plonk.py) by tests/test_oracle_c.py.
 *
 * `faithful` = 1 keeps the reference's per-element Fermat inversions (permutation.rs:276,
 * quotient_poly.rs:111); 0 uses batch inversion / the 8-entry Z_H table (a tuned CPU port).  Loops are
 * threaded with OpenMP in both modes (the reference threads only the permutation identity map,
 * quotient_poly.rs:243); results are identical.
 */
#include <omp.h>

#include "zkp_oracle_field.h"

static const fr_t *FR(const uint64_t *p) { return (const fr_t *)p; }

static inline fr_t F_add(fr_t a, fr_t b) { fr_t r; fr_add(&r, &a, &b); return r; }
static inline fr_t F_sub(fr_t a, fr_t b) { fr_t r; fr_sub(&r, &a, &b); return r; }
static inline fr_t F_mul(fr_t a, fr_t b) { fr_t r; fr_mul(&r, &a, &b); return r; }
static inline fr_t F_sqr(fr_t a) { fr_t r; fr_mul(&r, &a, &a); return r; }
static inline fr_t F_u64(uint64_t v) { fr_t r; fr_set_u64(&r, v); return r; }
static inline fr_t F_one(void) { fr_t r; memcpy(r.l, FR_ONE, 32); return r; }
static inline fr_t F_zero(void) { fr_t r; memset(&r, 0, sizeof r); return r; }
static inline fr_t F_inv(fr_t a) { fr_t r; fr_inv(&r, &a); return r; }

/* ------------------------------------------------------------------ permutation accumulator */
int oracle_perm_z(const uint64_t *wires, const uint64_t *sigmas, const uint64_t *roots, const uint64_t *beta_,
                  const uint64_t *gamma_, size_t n, uint64_t *out, int faithful, int nthreads) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    const fr_t beta = *FR(beta_), gamma = *FR(gamma_);
    fr_t ks[4] = {F_one(), F_u64(7), F_u64(13), F_u64(17)};
    fr_t *num = (fr_t *)malloc(n * sizeof(fr_t)), *den = (fr_t *)malloc(n * sizeof(fr_t));
    if (!num || !den) return -2;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (size_t i = 0; i < n; i++) {
        fr_t nu = F_one(), de = F_one();
        for (int j = 0; j < 4; j++) {
            fr_t w = FR(wires)[(size_t)j * n + i];
            fr_t f = F_add(F_add(w, F_mul(F_mul(beta, ks[j]), FR(roots)[i])), gamma);
            fr_t g = F_add(F_add(w, F_mul(beta, FR(sigmas)[(size_t)j * n + i])), gamma);
            nu = F_mul(nu, f);
            de = F_mul(de, g);
        }
        num[i] = nu;
        den[i] = de;
        if (faithful) num[i] = F_mul(nu, F_inv(de));
    }
    if (!faithful) {  /* Montgomery batch inversion of den, then num / den */
        fr_t *pre = (fr_t *)malloc(n * sizeof(fr_t));
        fr_t acc = F_one();
        for (size_t i = 0; i < n; i++) { pre[i] = acc; acc = F_mul(acc, den[i]); }
        fr_t inv = F_inv(acc);
        for (size_t i = n; i-- > 0;) {
            fr_t di = F_mul(inv, pre[i]);
            inv = F_mul(inv, den[i]);
            num[i] = F_mul(num[i], di);
        }
        free(pre);
    }
    fr_t state = F_one();
    fr_t *z = (fr_t *)out;
    for (size_t i = 0; i < n; i++) { z[i] = state; state = F_mul(state, num[i]); }
    free(num); free(den);
    return 0;
}

/* ------------------------------------------------------------------ gate widgets */
static inline fr_t delta4(fr_t f, fr_t one, fr_t two, fr_t three) {
    return F_mul(F_mul(f, F_sub(f, one)), F_mul(F_sub(f, two), F_sub(f, three)));
}
static inline fr_t times(fr_t x, unsigned k) {  /* small-constant multiple by double-and-add */
    fr_t acc = F_zero(), b = x;
    while (k) { if (k & 1) acc = F_add(acc, b); b = F_add(b, b); k >>= 1; }
    return acc;
}

/* cols: 0-3 wires a b c d | 4 z | 5 pi | 6 l1 | 7-17 selectors q_m q_l q_r q_o q_c q_4 q_arith q_range q_logic
 * q_fixed q_var | 18-21 sigma | 22 linear | 23 v_h_coset_8n.  All n8 evaluations on the coset. */
int oracle_quotient(const uint64_t *const *cols, const uint64_t *challenges, size_t n8, int faithful,
                    uint64_t *out, int nthreads) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    const fr_t *a_ = FR(cols[0]), *b_ = FR(cols[1]), *c_ = FR(cols[2]), *d_ = FR(cols[3]), *z_ = FR(cols[4]);
    const fr_t *pi_ = FR(cols[5]), *l1_ = FR(cols[6]);
    const fr_t *sel[11];
    for (int j = 0; j < 11; j++) sel[j] = FR(cols[7 + j]);
    const fr_t *sg[4] = {FR(cols[18]), FR(cols[19]), FR(cols[20]), FR(cols[21])};
    const fr_t *lin = FR(cols[22]), *vh = FR(cols[23]);
    const fr_t alpha = FR(challenges)[0], beta = FR(challenges)[1], gamma = FR(challenges)[2];
    const fr_t rs = FR(challenges)[3], ls = FR(challenges)[4], fs = FR(challenges)[5], vs = FR(challenges)[6];
    const fr_t one = F_one(), two = F_add(one, one), three = F_add(two, one);
    const fr_t bk1 = F_mul(beta, F_u64(7)), bk2 = F_mul(beta, F_u64(13)), bk3 = F_mul(beta, F_u64(17));
    fr_t D = F_mul(F_u64(10240), F_inv(F_u64(10241)));
    fr_neg(&D, &D);
    fr_t zh_inv[8];
    for (int j = 0; j < 8; j++) zh_inv[j] = F_inv(vh[j % n8]);
    fr_t *o = (fr_t *)out;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (size_t i = 0; i < n8; i++) {
        const size_t in = (i + 8) & (n8 - 1);
        const fr_t a = a_[i], b = b_[i], c = c_[i], d = d_[i], an = a_[in], bn = b_[in], dn = d_[in];
        const fr_t qc = sel[4][i];
        /* arithmetic + PI */
        fr_t t = F_mul(F_mul(a, b), sel[0][i]);
        t = F_add(t, F_mul(a, sel[1][i]));
        t = F_add(t, F_mul(b, sel[2][i]));
        t = F_add(t, F_mul(c, sel[3][i]));
        t = F_add(t, F_mul(d, sel[5][i]));
        t = F_add(t, qc);
        t = F_add(F_mul(t, sel[6][i]), pi_[i]);
        if (!fr_is_zero(&sel[7][i])) {  /* range */
            fr_t k = F_sqr(rs), k2 = F_sqr(k), k3 = F_mul(k2, k);
            fr_t s = delta4(F_sub(c, times(d, 4)), one, two, three);
            s = F_add(s, F_mul(delta4(F_sub(b, times(c, 4)), one, two, three), k));
            s = F_add(s, F_mul(delta4(F_sub(a, times(b, 4)), one, two, three), k2));
            s = F_add(s, F_mul(delta4(F_sub(dn, times(a, 4)), one, two, three), k3));
            t = F_add(t, F_mul(F_mul(s, rs), sel[7][i]));
        }
        if (!fr_is_zero(&sel[8][i])) {  /* logic */
            fr_t k = F_sqr(ls), k2 = F_sqr(k), k3 = F_mul(k2, k), k4 = F_mul(k3, k);
            fr_t A = F_sub(an, times(a, 4)), B = F_sub(bn, times(b, 4)), Dd = F_sub(dn, times(d, 4));
            fr_t ab = F_add(A, B);
            fr_t f = F_add(F_sub(times(c, 4), times(ab, 18)), times(one, 81));
            f = F_add(F_sub(F_add(F_mul(c, f), times(F_add(F_sqr(A), F_sqr(B)), 18)), times(ab, 81)), times(one, 83));
            f = F_mul(c, f);
            fr_t e = F_sub(times(F_add(ab, Dd), 3), F_add(f, f));
            fr_t bb = F_mul(qc, F_sub(times(Dd, 9), times(ab, 3)));
            fr_t s = F_mul(F_sub(c, F_mul(A, B)), k3);
            s = F_add(s, delta4(A, one, two, three));
            s = F_add(s, F_mul(delta4(B, one, two, three), k));
            s = F_add(s, F_mul(delta4(Dd, one, two, three), k2));
            s = F_add(s, F_mul(F_add(bb, e), k4));
            t = F_add(t, F_mul(F_mul(s, ls), sel[8][i]));
        }
        if (!fr_is_zero(&sel[9][i])) {  /* fixed-base */
            fr_t k = F_sqr(fs), k2 = F_sqr(k), k3 = F_mul(k2, k);
            fr_t xb = sel[1][i], yb = sel[2][i];
            fr_t bit = F_sub(dn, F_add(d, d));
            fr_t bitc = F_mul(F_mul(bit, F_sub(bit, one)), F_add(bit, one));
            fr_t ya = F_add(F_mul(F_sqr(bit), F_sub(yb, one)), one);
            fr_t xa = F_mul(bit, xb);
            fr_t xyc = F_mul(F_sub(F_mul(bit, qc), c), k);
            fr_t tt = F_mul(F_mul(F_mul(c, a), b), D);
            fr_t xacc = F_mul(F_sub(F_add(an, F_mul(an, tt)), F_add(F_mul(a, ya), F_mul(b, xa))), k2);
            fr_t yacc = F_mul(F_sub(F_sub(bn, F_mul(bn, tt)), F_add(F_mul(b, ya), F_mul(a, xa))), k3);
            fr_t s = F_add(F_add(bitc, xacc), F_add(yacc, xyc));
            t = F_add(t, F_mul(F_mul(s, fs), sel[9][i]));
        }
        if (!fr_is_zero(&sel[10][i])) {  /* variable-base */
            fr_t k = F_sqr(vs);
            fr_t y1x2 = F_mul(b, c), y1y2 = F_mul(b, d), x1x2 = F_mul(a, c);
            fr_t tt = F_mul(F_mul(D, dn), y1x2);
            fr_t xyc = F_sub(F_mul(a, d), dn);
            fr_t x3c = F_mul(F_sub(F_add(dn, y1x2), F_add(an, F_mul(an, tt))), k);
            fr_t y3c = F_mul(F_sub(F_add(y1y2, x1x2), F_sub(bn, F_mul(bn, tt))), F_sqr(k));
            t = F_add(t, F_mul(F_mul(F_add(F_add(xyc, x3c), y3c), vs), sel[10][i]));
        }
        {   /* permutation (quotient_poly.rs:245-261) */
            fr_t z = z_[i], zn = z_[in], x = lin[i];
            fr_t ag = F_add(a, gamma), bg = F_add(b, gamma), cg = F_add(c, gamma), dg = F_add(d, gamma);
            fr_t ident = F_mul(F_mul(F_add(ag, F_mul(beta, x)), F_add(bg, F_mul(bk1, x))),
                               F_mul(F_add(cg, F_mul(bk2, x)), F_add(dg, F_mul(bk3, x))));
            ident = F_mul(ident, z);
            fr_t copy = F_mul(F_mul(F_add(ag, F_mul(beta, sg[0][i])), F_add(bg, F_mul(beta, sg[1][i]))),
                              F_mul(F_add(cg, F_mul(beta, sg[2][i])), F_add(dg, F_mul(beta, sg[3][i]))));
            copy = F_mul(copy, zn);
            t = F_add(t, F_mul(F_sub(ident, copy), alpha));
            t = F_add(t, F_mul(F_sub(z, one), l1_[i]));
        }
        o[i] = F_mul(t, faithful ? F_inv(vh[i]) : zh_inv[i & 7]);
    }
    return 0;
}

/* ------------------------------------------------------------------ polynomial helpers */
int oracle_poly_eval(const uint64_t *p, size_t len, const uint64_t *point, uint64_t *out, int nthreads) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    const fr_t x = *FR(point);
    fr_t total = F_zero();
    if (len) {
        int T = nthreads;
        if ((size_t)T > len) T = (int)len;
        fr_t *part = (fr_t *)malloc(T * sizeof(fr_t));
#pragma omp parallel num_threads(T)
        {
            int t = omp_get_thread_num(), TT = omp_get_num_threads();
            size_t lo = len * t / TT, hi = len * (t + 1) / TT;
            fr_t s = F_zero();
            for (size_t i = hi; i-- > lo;) s = F_add(F_mul(s, x), FR(p)[i]);
            fr_t xl; fr_pow_u64(&xl, &x, lo);
            if (t < T) part[t] = F_mul(s, xl);
        }
        for (int t = 0; t < T; t++) total = F_add(total, part[t]);
        free(part);
    }
    memcpy(out, total.l, 32);
    return 0;
}

int oracle_poly_lincomb(const uint64_t *const *polys, const uint64_t *lens, const uint64_t *scalars, unsigned count,
                        uint64_t *out, size_t out_len, int nthreads) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    fr_t *o = (fr_t *)out;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (size_t i = 0; i < out_len; i++) {
        fr_t acc = F_zero();
        for (unsigned k = 0; k < count; k++)
            if (i < lens[k]) acc = F_add(acc, F_mul(FR(scalars)[k], FR(polys[k])[i]));
        o[i] = acc;
    }
    return 0;
}

/* writes len - 1 coefficients */
int oracle_poly_div_linear(const uint64_t *p, size_t len, const uint64_t *point, uint64_t *out) {
    const fr_t x = *FR(point);
    fr_t carry = F_zero();
    fr_t *q = (fr_t *)out;
    for (size_t i = len; i-- > 1;) {
        carry = F_add(FR(p)[i], F_mul(carry, x));
        q[i - 1] = carry;
    }
    return 0;
}

// BLS12-381 G1 (y^2 = x^3 + 4 over Fq) point arithmetic for the MSM kernels.
//
// Bucket accumulators use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2): a mixed addition costs 8M + 2S and never needs Z itself.  All edge cases
// (accumulator at infinity, equal points -> doubling, opposite points -> infinity, affine
// infinity encoded as x = y = 0) are handled exactly: commitments must be bit-identical
// to the reference's msm_curve_addition + Commitment::new (src/prover/proof.rs:450-454,
// 507-526) for every input, including repeated bases.
#pragma once
#include "arith.cuh"

namespace zkp {

struct alignas(16) g1_affine {
    fq_t x, y;
    ZKP_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static ZKP_HD g1_affine inf() { g1_affine r; r.x = fq_t::zero(); r.y = fq_t::zero(); return r; }
};

// window-table slot (msm.cu): x in bytes [0, 48), y in [64, 112) of a 128-byte entry
struct alignas(128) g1_tab {
    fq_t x; uint32_t pad0[4];
    fq_t y; uint32_t pad1[4];
};

struct alignas(16) g1_xyzz {
    fq_t x, y, zz, zzz;
    ZKP_HD bool is_inf() const { return zz.is_zero(); }
    static ZKP_HD g1_xyzz inf() {
        g1_xyzz r; r.x = fq_t::zero(); r.y = fq_t::zero(); r.zz = fq_t::zero(); r.zzz = fq_t::zero();
        return r;
    }
};

// acc <- 2 * q (q affine, finite)
ZKP_HD void xyzz_mdbl(g1_xyzz& acc, const g1_affine& q) {
    if (q.y.is_zero()) { acc = g1_xyzz::inf(); return; }
    fq_t U = dbl(q.y);
    fq_t V = sqr(U);
    fq_t W = U * V;
    fq_t S = q.x * V;
    fq_t X2 = sqr(q.x);
    fq_t M = dbl(X2) + X2;
    fq_t X3 = sqr(M) - dbl(S);
    acc.y = M * (S - X3) - W * q.y;
    acc.x = X3;
    acc.zz = V;
    acc.zzz = W;
}

// acc <- 2 * acc
ZKP_HD void xyzz_dbl(g1_xyzz& acc) {
    if (acc.is_inf()) return;
    if (acc.y.is_zero()) { acc = g1_xyzz::inf(); return; }
    fq_t U = dbl(acc.y);
    fq_t V = sqr(U);
    fq_t W = U * V;
    fq_t S = acc.x * V;
    fq_t X2 = sqr(acc.x);
    fq_t M = dbl(X2) + X2;
    fq_t X3 = sqr(M) - dbl(S);
    acc.y = M * (S - X3) - W * acc.y;
    acc.x = X3;
    acc.zz = V * acc.zz;
    acc.zzz = W * acc.zzz;
}

// acc <- acc + q (q affine; madd-2008-s)
ZKP_HD void xyzz_madd(g1_xyzz& acc, const g1_affine& q) {
    if (q.is_inf()) return;
    if (acc.is_inf()) {
        acc.x = q.x; acc.y = q.y; acc.zz = fq_t::one(); acc.zzz = fq_t::one();
        return;
    }
    fq_t P = q.x * acc.zz - acc.x;
    fq_t R = q.y * acc.zzz - acc.y;
    if (P.is_zero()) {
        if (R.is_zero()) xyzz_mdbl(acc, q);
        else acc = g1_xyzz::inf();
        return;
    }
    fq_t PP = sqr(P);
    fq_t PPP = P * PP;
    fq_t Q = acc.x * PP;
    fq_t X3 = sqr(R) - PPP - dbl(Q);
    acc.y = R * (Q - X3) - acc.y * PPP;
    acc.x = X3;
    acc.zz = acc.zz * PP;
    acc.zzz = acc.zzz * PPP;
}

// acc <- acc + b (add-2008-s)
ZKP_HD void xyzz_add(g1_xyzz& acc, const g1_xyzz& b) {
    if (b.is_inf()) return;
    if (acc.is_inf()) { acc = b; return; }
    fq_t U1 = acc.x * b.zz;
    fq_t S1 = acc.y * b.zzz;
    fq_t P = b.x * acc.zz - U1;
    fq_t R = b.y * acc.zzz - S1;
    if (P.is_zero()) {
        if (R.is_zero()) xyzz_dbl(acc);
        else acc = g1_xyzz::inf();
        return;
    }
    fq_t PP = sqr(P);
    fq_t PPP = P * PP;
    fq_t Q = U1 * PP;
    fq_t X3 = sqr(R) - PPP - dbl(Q);
    acc.y = R * (Q - X3) - S1 * PPP;
    acc.x = X3;
    acc.zz = acc.zz * b.zz * PP;
    acc.zzz = acc.zzz * b.zzz * PPP;
}

// One Fermat inversion: 1/ZZZ, then 1/ZZ = (ZZ/ZZZ)^2.
ZKP_HD g1_affine xyzz_to_affine(const g1_xyzz& a) {
    if (a.is_inf()) return g1_affine::inf();
    fq_t t = inverse(a.zzz);
    fq_t zinv = a.zz * t;
    g1_affine r;
    r.x = a.x * sqr(zinv);
    r.y = a.y * t;
    return r;
}

}  // namespace zkp

// Host side of the round driver, shared by the prover (create_proof.cu) and the verifier (verifier.cu):
// 64-bit-limb Montgomery fields over the device arithmetic's constants, the dusk encodings (canonical
// scalars, compressed G1, wide reduction), the Merlin transcript and the scalar side of the linearisation.
// References: src/prover.rs:99-105 (transcript labels), src/prover/linearization_poly.rs:75-105,136-225.
#pragma once
#include <stdint.h>
#include <string.h>

#include "common.cuh"
#include "host_inv.h"

namespace zkp {
namespace drv {

typedef unsigned __int128 u128;

// ------------------------------------------------------------------ host Montgomery fields
// 64-bit-limb CIOS over the same constants as the device arithmetic (arith.cuh parameter tables).
template <class P>
struct HostField {
    static constexpr int N = P::N / 2;
    struct el {
        uint64_t l[N];
        bool is_zero() const { uint64_t x = 0; for (int i = 0; i < N; i++) x |= l[i]; return x == 0; }
    };
    struct Consts {
        uint64_t p[N], one[N], r2[N], inv;
        Consts() {
            for (int i = 0; i < N; i++) {
                p[i] = (uint64_t)P::p(2 * i) | ((uint64_t)P::p(2 * i + 1) << 32);
                one[i] = (uint64_t)P::one(2 * i) | ((uint64_t)P::one(2 * i + 1) << 32);
                r2[i] = (uint64_t)P::r2(2 * i) | ((uint64_t)P::r2(2 * i + 1) << 32);
            }
            uint64_t x = 1;  // Newton: x <- x (2 - p0 x) doubles the correct low bits
            for (int i = 0; i < 7; i++) x *= 2 - p[0] * x;
            inv = 0 - x;
        }
    };
    static const Consts& C() { static const Consts c; return c; }

    static el zero() { el r; memset(r.l, 0, sizeof r.l); return r; }
    static el one() { el r; memcpy(r.l, C().one, sizeof r.l); return r; }
    static bool geq_p(const uint64_t* a) {
        const uint64_t* p = C().p;
        for (int i = N - 1; i >= 0; i--) { if (a[i] > p[i]) return true; if (a[i] < p[i]) return false; }
        return true;
    }
    static void sub_p(uint64_t* a) {
        const uint64_t* p = C().p;
        uint64_t br = 0;
        for (int i = 0; i < N; i++) { u128 d = (u128)a[i] - p[i] - br; a[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    }
    // a may be any value below 2^(64 N) (not necessarily reduced); b < p
    static el mul(const el& a, const el& b) {
        const Consts& c = C();
        uint64_t t[N + 2];
        memset(t, 0, sizeof t);
        for (int i = 0; i < N; i++) {
            uint64_t cy = 0;
            for (int j = 0; j < N; j++) { u128 s = (u128)a.l[j] * b.l[i] + t[j] + cy; t[j] = (uint64_t)s; cy = (uint64_t)(s >> 64); }
            u128 s = (u128)t[N] + cy; t[N] = (uint64_t)s; t[N + 1] = (uint64_t)(s >> 64);
            const uint64_t m = t[0] * c.inv;
            s = (u128)m * c.p[0] + t[0]; cy = (uint64_t)(s >> 64);
            for (int j = 1; j < N; j++) { s = (u128)m * c.p[j] + t[j] + cy; t[j - 1] = (uint64_t)s; cy = (uint64_t)(s >> 64); }
            s = (u128)t[N] + cy; t[N - 1] = (uint64_t)s; t[N] = t[N + 1] + (uint64_t)(s >> 64);
        }
        if (t[N] || geq_p(t)) sub_p(t);
        el r; memcpy(r.l, t, sizeof r.l); return r;
    }
    static el sqr(const el& a) { return mul(a, a); }
    static el add(const el& a, const el& b) {
        uint64_t t[N]; uint64_t cy = 0;
        for (int i = 0; i < N; i++) { u128 s = (u128)a.l[i] + b.l[i] + cy; t[i] = (uint64_t)s; cy = (uint64_t)(s >> 64); }
        if (cy || geq_p(t)) sub_p(t);
        el r; memcpy(r.l, t, sizeof r.l); return r;
    }
    static el neg(const el& a) {
        if (a.is_zero()) return a;
        const uint64_t* p = C().p;
        el r; uint64_t br = 0;
        for (int i = 0; i < N; i++) { u128 d = (u128)p[i] - a.l[i] - br; r.l[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
        return r;
    }
    static el sub(const el& a, const el& b) { return add(a, neg(b)); }
    static el dbl(const el& a) { return add(a, a); }
    static el to_mont(const el& raw) { el r2; memcpy(r2.l, C().r2, sizeof r2.l); return mul(raw, r2); }
    static el from_mont(const el& a) { el o = zero(); o.l[0] = 1; return mul(a, o); }
    static el from_u64(uint64_t v) { el r = zero(); r.l[0] = v; return to_mont(r); }
    static el pow(el base, uint64_t e) {
        el acc = one();
        while (e) { if (e & 1) acc = mul(acc, base); base = sqr(base); e >>= 1; }
        return acc;
    }
    static el inv(const el& a) {  // Montgomery form in and out; 0 -> 0.  Binary GCD (host_inv.h) on a R, then R^3 R^-1
        static const el R3 = [] { el r2; memcpy(r2.l, C().r2, sizeof r2.l); return mul(r2, r2); }();
        el x;
        hostinv::inv_mod<N>(a.l, C().p, x.l);
        return mul(x, R3);
    }
};

typedef HostField<FrParams> F;
typedef HostField<FqParams> Q;
typedef F::el fr;

static inline fr fr_load(const uint64_t* p) { fr r; memcpy(r.l, p, 32); return r; }
static inline void fr_store(uint64_t* p, const fr& a) { memcpy(p, a.l, 32); }

// 32-byte little-endian canonical encoding (TranscriptProtocol::append_scalar, proof wire format)
static void fr_bytes(const fr& mont, uint8_t out[32]) {
    const fr c = F::from_mont(mont);
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(c.l[i] >> (8 * b));
}

// 64 little-endian bytes reduced mod r, returned in Montgomery form (challenge_scalar):
// v = lo + 2^256 hi; mul(x, R^2) maps any 256-bit x to x R mod r, and R^2 is also the
// Montgomery form of 2^256
static fr fr_from_wide(const uint8_t b[64]) {
    fr lo = F::zero(), hi = F::zero();
    for (int i = 0; i < 32; i++) {
        lo.l[i / 8] |= (uint64_t)b[i] << (8 * (i % 8));
        hi.l[i / 8] |= (uint64_t)b[32 + i] << (8 * (i % 8));
    }
    fr r2; memcpy(r2.l, F::C().r2, 32);
    return F::add(F::mul(lo, r2), F::mul(F::mul(hi, r2), r2));
}

// 48-byte compressed G1 (big-endian x; bit 7 compressed, bit 6 infinity, bit 5 = y is the larger root)
static void g1_compress(const uint64_t xy[12], uint8_t out[48]) {
    uint64_t any = 0;
    for (int i = 0; i < 12; i++) any |= xy[i];
    memset(out, 0, 48);
    if (!any) { out[0] = 0xC0; return; }
    Q::el x, y;
    memcpy(x.l, xy, 48); memcpy(y.l, xy + 6, 48);
    const Q::el xc = Q::from_mont(x), yc = Q::from_mont(y);
    for (int i = 0; i < 6; i++)
        for (int b = 0; b < 8; b++) out[47 - (8 * i + b)] = (uint8_t)(xc.l[i] >> (8 * b));
    out[0] |= 0x80;
    // y > p - y  (canonical integers)
    const uint64_t* p = Q::C().p;
    uint64_t ny[6]; uint64_t br = 0;
    for (int i = 0; i < 6; i++) { u128 d = (u128)p[i] - yc.l[i] - br; ny[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    bool larger = false;
    for (int i = 5; i >= 0; i--) {
        if (yc.l[i] > ny[i]) { larger = true; break; }
        if (yc.l[i] < ny[i]) break;
    }
    if (larger) out[0] |= 0x20;
}

// ------------------------------------------------------------------ Merlin transcript
// STROBE-128/1600 restricted to the operations Merlin uses (meta-AD, AD, PRF).
struct Transcript {
    static constexpr int RATE = 166;
    uint8_t st[200];
    uint8_t pos, pos_begin, cur_flags;

    void load(const uint8_t in[203]) { memcpy(st, in, 200); pos = in[200]; pos_begin = in[201]; cur_flags = in[202]; }
    void save(uint8_t out[203]) const { memcpy(out, st, 200); out[200] = pos; out[201] = pos_begin; out[202] = cur_flags; }
    void run_f() {
        st[pos] ^= pos_begin;
        st[pos + 1] ^= 0x04;
        st[RATE + 1] ^= 0x80;
        uint64_t w[25];
        memcpy(w, st, 200);
        zkp_keccak_f1600(w);
        memcpy(st, w, 200);
        pos = 0; pos_begin = 0;
    }
    void absorb(const uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) { st[pos++] ^= d[i]; if (pos == RATE) run_f(); }
    }
    void squeeze(uint8_t* d, size_t n) {
        for (size_t i = 0; i < n; i++) { d[i] = st[pos]; st[pos++] = 0; if (pos == RATE) run_f(); }
    }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;
        const uint8_t old = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        const uint8_t hdr[2] = {old, flags};
        absorb(hdr, 2);
        if ((flags & (4 | 32)) && pos != 0) run_f();
    }
    void meta_ad(const void* d, size_t n, bool more) { begin_op(16 | 2, more); absorb((const uint8_t*)d, n); }
    void ad(const void* d, size_t n) { begin_op(2, false); absorb((const uint8_t*)d, n); }
    void append_message(const char* label, const uint8_t* msg, uint32_t n) {
        meta_ad(label, strlen(label), false);
        const uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        meta_ad(len, 4, true);
        ad(msg, n);
    }
    void challenge_bytes(const char* label, uint8_t* out, uint32_t n) {
        meta_ad(label, strlen(label), false);
        const uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
        meta_ad(len, 4, true);
        begin_op(1 | 2 | 4, false);
        squeeze(out, n);
    }
    // TranscriptProtocol (dusk encodings)
    void append_scalar(const char* label, const fr& mont) { uint8_t b[32]; fr_bytes(mont, b); append_message(label, b, 32); }
    void append_commitment(const char* label, const uint64_t xy[12]) { uint8_t b[48]; g1_compress(xy, b); append_message(label, b, 48); }
    fr challenge_scalar(const char* label) { uint8_t b[64]; challenge_bytes(label, b, 64); return fr_from_wide(b); }
};

// ------------------------------------------------------------------ linearisation scalars
// widget.linearize / permutation.linearize on the opened evaluations (Montgomery form throughout).
struct Lin {
    static fr c(uint64_t v) { return F::from_u64(v); }
    static fr m(const fr& a, const fr& b) { return F::mul(a, b); }
    static fr a(const fr& x, const fr& y) { return F::add(x, y); }
    static fr s(const fr& x, const fr& y) { return F::sub(x, y); }
    static fr delta(const fr& f) {
        const fr one = F::one();
        const fr f1 = s(f, one), f2 = s(f1, one), f3 = s(f2, one);
        return m(m(m(f, f1), f2), f3);
    }
    static fr x4(const fr& v) { return F::dbl(F::dbl(v)); }
    static fr edwards_d() {  // JubJub d = -(10240 / 10241)
        static const fr d = F::neg(m(c(10240), F::inv(c(10241))));
        return d;
    }
    static fr range_term(const fr& sep, const fr& A, const fr& B, const fr& C, const fr& D, const fr& Dn) {
        const fr k = m(sep, sep), k2 = m(k, k), k3 = m(k2, k);
        fr r = delta(s(C, x4(D)));
        r = a(r, m(delta(s(B, x4(C))), k));
        r = a(r, m(delta(s(A, x4(B))), k2));
        r = a(r, m(delta(s(Dn, x4(A))), k3));
        return m(r, sep);
    }
    static fr logic_term(const fr& sep, const fr& wa, const fr& an, const fr& wb, const fr& bn, const fr& wc,
                         const fr& wd, const fr& dn, const fr& qc) {
        const fr k = m(sep, sep), k2 = m(k, k), k3 = m(k2, k), k4 = m(k3, k);
        const fr A = s(an, x4(wa)), B = s(bn, x4(wb)), D = s(dn, x4(wd));
        const fr AB = a(A, B);
        // f = c (c (4c - 18(A+B) + 81) + 18(A^2 + B^2) - 81(A+B) + 83)
        fr inner = a(s(x4(wc), m(c(18), AB)), c(81));
        inner = m(wc, inner);
        inner = a(inner, m(c(18), a(m(A, A), m(B, B))));
        inner = s(inner, m(c(81), AB));
        inner = a(inner, c(83));
        const fr f = m(wc, inner);
        const fr e = s(m(c(3), a(AB, D)), F::dbl(f));
        const fr bb = m(qc, s(m(c(9), D), m(c(3), AB)));
        fr r = m(s(wc, m(A, B)), k3);
        r = a(r, delta(A));
        r = a(r, m(delta(B), k));
        r = a(r, m(delta(D), k2));
        r = a(r, m(a(bb, e), k4));
        return m(r, sep);
    }
    static fr fixed_base_term(const fr& sep, const fr& wa, const fr& an, const fr& wb, const fr& bn, const fr& wc,
                              const fr& wd, const fr& dn, const fr& ql, const fr& qr, const fr& qc) {
        const fr one = F::one();
        const fr k = m(sep, sep), k2 = m(k, k), k3 = m(k2, k);
        const fr bit = s(dn, F::dbl(wd));
        const fr y_alpha = a(m(m(bit, bit), s(qr, one)), one);
        const fr x_alpha = m(bit, ql);
        const fr t = m(m(m(wc, wa), wb), edwards_d());
        const fr x_acc = m(s(a(an, m(an, t)), a(m(wa, y_alpha), m(wb, x_alpha))), k2);
        const fr y_acc = m(s(s(bn, m(bn, t)), a(m(wb, y_alpha), m(wa, x_alpha))), k3);
        fr r = m(m(bit, s(bit, one)), a(bit, one));
        r = a(r, x_acc);
        r = a(r, y_acc);
        r = a(r, m(s(m(bit, qc), wc), k));
        return m(r, sep);
    }
    static fr var_base_term(const fr& sep, const fr& wa, const fr& an, const fr& wb, const fr& bn, const fr& wc,
                            const fr& wd, const fr& dn) {
        const fr k = m(sep, sep);
        const fr y1x2 = m(wb, wc), y1y2 = m(wb, wd), x1x2 = m(wa, wc);
        const fr t = m(m(edwards_d(), dn), y1x2);
        const fr x3c = m(s(a(dn, y1x2), a(an, m(an, t))), k);
        const fr y3c = m(s(a(y1y2, x1x2), s(bn, m(bn, t))), m(k, k));
        return m(a(a(s(m(wa, wd), dn), x3c), y3c), sep);
    }
};

// r(X) = sum_j sc[j] * poly_j(X) over q_m q_l q_r q_o q_4 q_c q_range q_logic q_fixed q_var z s_sigma_4.
// ch = alpha beta gamma range logic fixed var z_challenge; e = the first 15 entries of `Evaluations`.
static void linearization_scalars(uint64_t n, const fr ch[8], const fr e[15], fr sc[12]) {
    const fr &alpha = ch[0], &beta = ch[1], &gamma = ch[2], &zc = ch[7];
    const fr &a = e[0], &b = e[1], &c = e[2], &d = e[3], &an = e[4], &bn = e[5], &dn = e[6], &s1 = e[7], &s2 = e[8],
             &s3 = e[9], &qarith = e[10], &qc = e[11], &ql = e[12], &qr = e[13], &pe = e[14];
    sc[0] = F::mul(F::mul(a, b), qarith);
    sc[1] = F::mul(a, qarith);
    sc[2] = F::mul(b, qarith);
    sc[3] = F::mul(c, qarith);
    sc[4] = F::mul(d, qarith);
    sc[5] = qarith;
    sc[6] = Lin::range_term(ch[3], a, b, c, d, dn);
    sc[7] = Lin::logic_term(ch[4], a, an, b, bn, c, d, dn, qc);
    sc[8] = Lin::fixed_base_term(ch[5], a, an, b, bn, c, d, dn, ql, qr, qc);
    sc[9] = Lin::var_base_term(ch[6], a, an, b, bn, c, d, dn);
    // permutation part (pinned by the verifier identity, src/prover/proof.rs:386-440)
    const fr one = F::one();
    const fr zh = F::sub(F::pow(zc, n), one);
    const fr l1 = F::mul(zh, F::inv(F::mul(F::from_u64(n), F::sub(zc, one))));
    const fr bz = F::mul(beta, zc);
    fr x = F::add(F::add(a, bz), gamma);
    x = F::mul(x, F::add(F::add(b, F::mul(bz, F::from_u64(7))), gamma));
    x = F::mul(x, F::add(F::add(c, F::mul(bz, F::from_u64(13))), gamma));
    x = F::mul(x, F::add(F::add(d, F::mul(bz, F::from_u64(17))), gamma));
    x = F::mul(x, alpha);
    fr y = F::add(F::add(a, F::mul(beta, s1)), gamma);
    y = F::mul(y, F::add(F::add(b, F::mul(beta, s2)), gamma));
    y = F::mul(y, F::add(F::add(c, F::mul(beta, s3)), gamma));
    y = F::mul(F::mul(F::mul(y, beta), pe), alpha);
    sc[10] = F::add(x, F::mul(l1, F::sqr(alpha)));
    sc[11] = F::neg(y);
}

}  // namespace drv
}  // namespace zkp

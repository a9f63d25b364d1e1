// Verifier glue on the product side: Proof::verify (src/prover/proof.rs:70-383),
// Verifier::verify (src/verifier.rs:46-81), batch_check (src/commitment_scheme.rs:24-66) and the G2 half of
// the opening key that PlonkParams::verification_key() hands over (src/key.rs:320).  Constant-size host
// work -- one transcript replay, a few dozen G1 scalar multiplications and ONE two-pairing product -- so it
// is plain host C++ over the 64-bit-limb fields of host_driver.h; nothing here runs on the GPU and nothing
// is on the proving path.
//
// The pairing crate (ec-pairing) is absent from the reference tree; this is the ate pairing on the public
// BLS12-381 parameters: Fq2 = Fq[u] / (u^2 + 1), G2 on y^2 = x^3 + 4 (u + 1), Fq12 = Fq[w] / (w^12 - 2 w^6 + 2)
// with u = w^6 - 1, Miller loop over |x| = 0xd201000000010000, final exponentiation (p^12 - 1) / r split as
// (p^6 - 1) (conjugate over inverse) times (p^6 + 1) / r.  G2 arithmetic stays in Fq2: a point (x, y) of the
// twist maps to (x w^-2, y w^-3), so the slope of a tangent / chord through twisted points is m w^-1 with m
// the slope over Fq2, and the line through R evaluated at P = (xP, yP) in G1 is
//     -yP + (m xP) w^-1 + (y - m x) w^-3.
#include <string.h>

#include <vector>

#include "host_driver.h"

namespace zkp {
namespace vf {

using drv::F;
using drv::Q;
using drv::fr;
typedef Q::el fq;

static fq fq_raw(const uint64_t w[6]) { fq r; memcpy(r.l, w, 48); return Q::to_mont(r); }
static bool fq_eq(const fq& a, const fq& b) { return memcmp(a.l, b.l, 48) == 0; }

// ------------------------------------------------------------------ G1 (XYZZ on the host)
struct P1 {
    fq x, y, zz, zzz;
    bool inf() const { return zz.is_zero(); }
};
static P1 p1_inf() { P1 p; p.x = p.y = p.zz = p.zzz = Q::zero(); return p; }
static P1 p1_affine(const uint64_t xy[12]) {
    uint64_t any = 0;
    for (int i = 0; i < 12; i++) any |= xy[i];
    if (!any) return p1_inf();
    P1 p;
    memcpy(p.x.l, xy, 48); memcpy(p.y.l, xy + 6, 48);
    p.zz = p.zzz = Q::one();
    return p;
}
static void p1_dbl(P1& a) {
    if (a.inf()) return;
    if (a.y.is_zero()) { a = p1_inf(); return; }
    const fq U = Q::dbl(a.y), V = Q::sqr(U), W = Q::mul(U, V), S = Q::mul(a.x, V), X2 = Q::sqr(a.x);
    const fq M = Q::add(Q::dbl(X2), X2);
    const fq X3 = Q::sub(Q::sqr(M), Q::dbl(S));
    a.y = Q::sub(Q::mul(M, Q::sub(S, X3)), Q::mul(W, a.y));
    a.x = X3;
    a.zz = Q::mul(V, a.zz);
    a.zzz = Q::mul(W, a.zzz);
}
static void p1_add(P1& a, const P1& b) {
    if (b.inf()) return;
    if (a.inf()) { a = b; return; }
    const fq U1 = Q::mul(a.x, b.zz), S1 = Q::mul(a.y, b.zzz);
    const fq Pp = Q::sub(Q::mul(b.x, a.zz), U1), R = Q::sub(Q::mul(b.y, a.zzz), S1);
    if (Pp.is_zero()) {
        if (R.is_zero()) p1_dbl(a); else a = p1_inf();
        return;
    }
    const fq PP = Q::sqr(Pp), PPP = Q::mul(Pp, PP), Qq = Q::mul(U1, PP);
    const fq X3 = Q::sub(Q::sub(Q::sqr(R), PPP), Q::dbl(Qq));
    a.y = Q::sub(Q::mul(R, Q::sub(Qq, X3)), Q::mul(S1, PPP));
    a.x = X3;
    a.zz = Q::mul(Q::mul(a.zz, b.zz), PP);
    a.zzz = Q::mul(Q::mul(a.zzz, b.zzz), PPP);
}
static P1 p1_neg(P1 a) { a.y = Q::neg(a.y); return a; }
// s * p, s a Montgomery-form scalar
static P1 p1_mul(const P1& p, const fr& s) {
    const fr c = F::from_mont(s);
    P1 acc = p1_inf();
    for (int i = 254; i >= 0; i--) {
        p1_dbl(acc);
        if ((c.l[i >> 6] >> (i & 63)) & 1) p1_add(acc, p);
    }
    return acc;
}
// k * p for a raw 256-bit little-endian integer k (not reduced: the subgroup check multiplies by r itself)
static P1 p1_mul_raw(const P1& p, const uint64_t k[4]) {
    P1 acc = p1_inf();
    for (int i = 255; i >= 0; i--) {
        p1_dbl(acc);
        if ((k[i >> 6] >> (i & 63)) & 1) p1_add(acc, p);
    }
    return acc;
}
// a^e for a little-endian multi-word exponent
static fq fq_pow_words(const fq& a, const uint64_t* e, int words) {
    fq acc = Q::one();
    for (int i = 64 * words - 1; i >= 0; i--) {
        acc = Q::sqr(acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = Q::mul(acc, a);
    }
    return acc;
}
// 48-byte compressed G1 (zcash / dusk G1Affine::from_bytes) -> affine Montgomery; false on ANY malformed input:
// missing compression flag, x >= p, stray bits with the infinity flag, x not on the curve, point outside the
// r-torsion (BLS12-381 G1 has a cofactor)
static bool g1_decompress(const uint8_t in[48], uint64_t out_xy[12]) {
    memset(out_xy, 0, 96);
    if (!(in[0] & 0x80)) return false;
    if (in[0] & 0x40) {
        if (in[0] & 0x3f) return false;
        for (int i = 1; i < 48; i++) if (in[i]) return false;
        return true;
    }
    fq xr;
    memset(xr.l, 0, 48);
    for (int i = 0; i < 48; i++) {
        const uint8_t b = i == 0 ? (uint8_t)(in[0] & 0x1f) : in[i];
        xr.l[(47 - i) / 8] |= (uint64_t)b << (8 * ((47 - i) % 8));
    }
    if (Q::geq_p(xr.l)) return false;
    const fq x = Q::to_mont(xr);
    const fq y2 = Q::add(Q::mul(Q::sqr(x), x), Q::from_u64(4));
    // p = 3 mod 4: a square root of y2, if there is one, is y2^((p + 1) / 4)
    uint64_t e[6];
    {
        const uint64_t* p = Q::C().p;
        unsigned __int128 c = 1;
        uint64_t t[6];
        for (int i = 0; i < 6; i++) { c += p[i]; t[i] = (uint64_t)c; c >>= 64; }   // p + 1 (no carry out: p < 2^381)
        for (int i = 0; i < 6; i++) e[i] = (t[i] >> 2) | (i < 5 ? t[i + 1] << 62 : 0);
    }
    fq y = fq_pow_words(y2, e, 6);
    if (!fq_eq(Q::sqr(y), y2)) return false;
    // the flag says whether y is the lexicographically larger root
    const fq yc = Q::from_mont(y), ny = Q::from_mont(Q::neg(y));
    bool larger = false;
    for (int i = 5; i >= 0; i--) {
        if (yc.l[i] > ny.l[i]) { larger = true; break; }
        if (yc.l[i] < ny.l[i]) break;
    }
    if (larger != ((in[0] & 0x20) != 0)) y = Q::neg(y);
    P1 pt;
    pt.x = x; pt.y = y; pt.zz = pt.zzz = Q::one();
    if (!p1_mul_raw(pt, F::C().p).inf()) return false;      // [r] P must be the identity
    memcpy(out_xy, x.l, 48); memcpy(out_xy + 6, y.l, 48);
    return true;
}

static void p1_to_affine(const P1& a, fq* x, fq* y, bool* inf) {
    *inf = a.inf();
    if (*inf) { *x = *y = Q::zero(); return; }
    const fq t = Q::inv(a.zzz), zi = Q::mul(a.zz, t);
    *x = Q::mul(a.x, Q::sqr(zi));
    *y = Q::mul(a.y, t);
}
static P1 p1_generator() {
    static const uint64_t gx[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                                   0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
    static const uint64_t gy[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                                   0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
    P1 p;
    p.x = fq_raw(gx); p.y = fq_raw(gy); p.zz = p.zzz = Q::one();
    return p;
}

// ------------------------------------------------------------------ Fq2 and G2 (affine)
struct fq2 { fq a, b; };   // a + b u
static fq2 f2_add(const fq2& x, const fq2& y) { return {Q::add(x.a, y.a), Q::add(x.b, y.b)}; }
static fq2 f2_sub(const fq2& x, const fq2& y) { return {Q::sub(x.a, y.a), Q::sub(x.b, y.b)}; }
static fq2 f2_mul(const fq2& x, const fq2& y) {
    const fq aa = Q::mul(x.a, y.a), bb = Q::mul(x.b, y.b);
    const fq cross = Q::mul(Q::add(x.a, x.b), Q::add(y.a, y.b));
    return {Q::sub(aa, bb), Q::sub(Q::sub(cross, aa), bb)};
}
static fq2 f2_inv(const fq2& x) {
    const fq d = Q::inv(Q::add(Q::sqr(x.a), Q::sqr(x.b)));
    return {Q::mul(x.a, d), Q::neg(Q::mul(x.b, d))};
}
static bool f2_zero(const fq2& x) { return x.a.is_zero() && x.b.is_zero(); }
static bool f2_eq(const fq2& x, const fq2& y) { return fq_eq(x.a, y.a) && fq_eq(x.b, y.b); }

struct P2 { fq2 x, y; bool inf; };
static P2 p2_inf() { P2 p; p.x = p.y = {Q::zero(), Q::zero()}; p.inf = true; return p; }
// slope of the tangent at p / the chord through p, q; false when the sum is the point at infinity
static bool p2_slope(const P2& p, const P2& q, fq2* m) {
    if (f2_eq(p.x, q.x)) {
        if (!f2_eq(p.y, q.y) || f2_zero(p.y)) return false;
        const fq2 xx = f2_mul(p.x, p.x);
        *m = f2_mul(f2_add(f2_add(xx, xx), xx), f2_inv(f2_add(p.y, p.y)));
    } else {
        *m = f2_mul(f2_sub(q.y, p.y), f2_inv(f2_sub(q.x, p.x)));
    }
    return true;
}
static P2 p2_add_with(const P2& p, const P2& q, const fq2& m) {
    P2 r;
    r.inf = false;
    r.x = f2_sub(f2_sub(f2_mul(m, m), p.x), q.x);
    r.y = f2_sub(f2_mul(m, f2_sub(p.x, r.x)), p.y);
    return r;
}
static P2 p2_add(const P2& p, const P2& q) {
    if (p.inf) return q;
    if (q.inf) return p;
    fq2 m;
    if (!p2_slope(p, q, &m)) return p2_inf();
    return p2_add_with(p, q, m);
}
static P2 p2_mul(const P2& p, const fr& s) {
    const fr c = F::from_mont(s);
    P2 acc = p2_inf();
    for (int i = 254; i >= 0; i--) {
        acc = p2_add(acc, acc);
        if ((c.l[i >> 6] >> (i & 63)) & 1) acc = p2_add(acc, p);
    }
    return acc;
}
static P2 p2_generator() {
    static const uint64_t c[4][6] = {
        {0xd48056c8c121bdb8ULL, 0x0bac0326a805bbefULL, 0xb4510b647ae3d177ULL, 0xc6e47ad4fa403b02ULL, 0x260805272dc51051ULL, 0x024aa2b2f08f0a91ULL},
        {0xe5ac7d055d042b7eULL, 0x334cf11213945d57ULL, 0xb5da61bbdc7f5049ULL, 0x596bd0d09920b61aULL, 0x7dacd3a088274f65ULL, 0x13e02b6052719f60ULL},
        {0xe193548608b82801ULL, 0x923ac9cc3baca289ULL, 0x6d429a695160d12cULL, 0xadfd9baa8cbdd3a7ULL, 0x8cc9cdc6da2e351aULL, 0x0ce5d527727d6e11ULL},
        {0xaaa9075ff05f79beULL, 0x3f370d275cec1da1ULL, 0x267492ab572e99abULL, 0xcb3e287e85a763afULL, 0x32acd2b02bc28b99ULL, 0x0606c4a02ea734ccULL}};
    P2 g;
    g.inf = false;
    g.x = {fq_raw(c[0]), fq_raw(c[1])};
    g.y = {fq_raw(c[2]), fq_raw(c[3])};
    return g;
}
static P2 p2_load(const uint64_t w[24]) {
    uint64_t any = 0;
    for (int i = 0; i < 24; i++) any |= w[i];
    if (!any) return p2_inf();
    P2 p;
    p.inf = false;
    memcpy(p.x.a.l, w, 48); memcpy(p.x.b.l, w + 6, 48); memcpy(p.y.a.l, w + 12, 48); memcpy(p.y.b.l, w + 18, 48);
    return p;
}
static void p2_store(const P2& p, uint64_t w[24]) {
    memset(w, 0, 24 * sizeof(uint64_t));
    if (p.inf) return;
    memcpy(w, p.x.a.l, 48); memcpy(w + 6, p.x.b.l, 48); memcpy(w + 12, p.y.a.l, 48); memcpy(w + 18, p.y.b.l, 48);
}
static bool p2_on_curve(const P2& p) {
    if (p.inf) return true;
    const fq four = Q::from_u64(4);
    const fq2 b2 = {four, four};
    return f2_eq(f2_mul(p.y, p.y), f2_add(f2_mul(f2_mul(p.x, p.x), p.x), b2));
}

// ------------------------------------------------------------------ Fq12 = Fq[w] / (w^12 - 2 w^6 + 2)
struct f12 { fq c[12]; };
static f12 f12_zero() { f12 r; for (int i = 0; i < 12; i++) r.c[i] = Q::zero(); return r; }
static f12 f12_one() { f12 r = f12_zero(); r.c[0] = Q::one(); return r; }
static bool f12_is_one(const f12& a) {
    if (!fq_eq(a.c[0], Q::one())) return false;
    for (int i = 1; i < 12; i++) if (!a.c[i].is_zero()) return false;
    return true;
}
static f12 f12_mul(const f12& a, const f12& b) {
    fq t[23];
    for (int i = 0; i < 23; i++) t[i] = Q::zero();
    for (int i = 0; i < 12; i++) {
        if (a.c[i].is_zero()) continue;
        for (int j = 0; j < 12; j++) {
            if (b.c[j].is_zero()) continue;
            t[i + j] = Q::add(t[i + j], Q::mul(a.c[i], b.c[j]));
        }
    }
    for (int i = 22; i >= 12; i--) {   // w^i = 2 w^(i-6) - 2 w^(i-12)
        if (t[i].is_zero()) continue;
        const fq two = Q::dbl(t[i]);
        t[i - 6] = Q::add(t[i - 6], two);
        t[i - 12] = Q::sub(t[i - 12], two);
    }
    f12 r;
    for (int i = 0; i < 12; i++) r.c[i] = t[i];
    return r;
}
static f12 f12_conj(f12 a) {   // x^(p^6): w -> -w
    for (int i = 1; i < 12; i += 2) a.c[i] = Q::neg(a.c[i]);
    return a;
}
// inverse by extended Euclid in Fq[w] against the modulus polynomial (used a handful of times per check)
static int pdeg(const std::vector<fq>& p) { int d = (int)p.size() - 1; while (d > 0 && p[d].is_zero()) d--; return d; }
static f12 f12_inv(const f12& a) {
    std::vector<fq> lm(13, Q::zero()), hm(13, Q::zero()), low(13, Q::zero()), high(13, Q::zero());
    lm[0] = Q::one();
    for (int i = 0; i < 12; i++) low[i] = a.c[i];
    high[0] = Q::from_u64(2); high[6] = Q::neg(Q::from_u64(2)); high[12] = Q::one();
    while (pdeg(low) > 0) {
        // r = high / low (rounded polynomial division)
        const int dl = pdeg(low);
        std::vector<fq> temp(high), r(13, Q::zero());
        const fq il = Q::inv(low[dl]);
        for (int i = pdeg(high) - dl; i >= 0; i--) {
            r[i] = Q::add(r[i], Q::mul(temp[dl + i], il));
            for (int c = 0; c <= dl; c++) temp[c + i] = Q::sub(temp[c + i], Q::mul(r[i], low[c]));
        }
        std::vector<fq> nm(hm), nw(high);
        for (int i = 0; i < 13; i++)
            for (int j = 0; j < 13 - i; j++) {
                if (r[j].is_zero()) continue;
                nm[i + j] = Q::sub(nm[i + j], Q::mul(lm[i], r[j]));
                nw[i + j] = Q::sub(nw[i + j], Q::mul(low[i], r[j]));
            }
        hm = lm; high = low; lm = nm; low = nw;
    }
    const fq i0 = Q::inv(low[0]);
    f12 out;
    for (int i = 0; i < 12; i++) out.c[i] = Q::mul(lm[i], i0);
    return out;
}
// (a + b u) * k for a fixed Fq12 constant k, with u = w^6 - 1: (a - b) k + b (w^6 k)
struct Emb { f12 k, kw6; };
static Emb emb_of(const f12& k) {
    f12 w6 = f12_zero();
    w6.c[6] = Q::one();
    return {k, f12_mul(w6, k)};
}
static f12 emb_mul(const fq2& x, const Emb& e) {
    const fq d = Q::sub(x.a, x.b);
    f12 r;
    for (int i = 0; i < 12; i++) r.c[i] = Q::add(Q::mul(d, e.k.c[i]), Q::mul(x.b, e.kw6.c[i]));
    return r;
}

struct PairingConsts {
    Emb w1i, w3i;   // w^-1, w^-3
    PairingConsts() {
        f12 w = f12_zero();
        w.c[1] = Q::one();
        const f12 wi = f12_inv(w);
        w1i = emb_of(wi);
        w3i = emb_of(f12_mul(f12_mul(wi, wi), wi));
    }
};
static const PairingConsts& PC() { static const PairingConsts c; return c; }

// line through R (slope m over Fq2) evaluated at the G1 point (xp, yp)
static f12 line_value(const P2& R, const fq2& m, const fq& xp, const fq& yp) {
    const fq2 mx = {Q::mul(m.a, xp), Q::mul(m.b, xp)};
    f12 l = emb_mul(mx, PC().w1i);
    const f12 c3 = emb_mul(f2_sub(R.y, f2_mul(m, R.x)), PC().w3i);
    for (int i = 0; i < 12; i++) l.c[i] = Q::add(l.c[i], c3.c[i]);
    l.c[0] = Q::sub(l.c[0], yp);
    return l;
}
// vertical line through R at xp: xp - x_R w^-2 (chord through R and -R)
static f12 vertical_value(const P2& R, const fq& xp) {
    f12 wi2 = f12_mul(PC().w1i.k, PC().w1i.k);
    f12 l = emb_mul(R.x, emb_of(wi2));
    for (int i = 0; i < 12; i++) l.c[i] = Q::neg(l.c[i]);
    l.c[0] = Q::add(l.c[0], xp);
    return l;
}

static f12 miller_loop(const P2& Qp, const fq& xp, const fq& yp, bool p_inf) {
    if (Qp.inf || p_inf) return f12_one();
    static const uint64_t ATE = 0xd201000000010000ULL;
    P2 R = Qp;
    f12 f = f12_one();
    for (int i = 62; i >= 0; i--) {
        fq2 m;
        if (p2_slope(R, R, &m)) {
            f = f12_mul(f12_mul(f, f), line_value(R, m, xp, yp));
            R = p2_add_with(R, R, m);
        } else {
            f = f12_mul(f12_mul(f, f), vertical_value(R, xp));
            R = p2_inf();
        }
        if ((ATE >> i) & 1) {
            if (R.inf) { R = Qp; continue; }
            if (p2_slope(R, Qp, &m)) {
                f = f12_mul(f, line_value(R, m, xp, yp));
                R = p2_add_with(R, Qp, m);
            } else {
                f = f12_mul(f, vertical_value(R, xp));
                R = p2_inf();
            }
        }
    }
    return f;
}

// ---- final exponentiation through the Frobenius endomorphism ------------------------------------------
// x^p for x = sum c_i w^i with c_i in Fq is sum c_i (w^p)^i: a 12 x 12 matrix of constants, computed once from
// w^p (one 381-bit exponentiation).  Easy part f^((p^6 - 1)(p^2 + 1)), then for the hard part the identity
//     3 (p^4 - p^2 + 1) / r = (x - 1)^2 (x + p) (x^2 + p^2 - 1) + 3            (x the BLS parameter, negative)
// (Hayashida, Hayasaka, Teruya 2020): five exponentiations by |x| (64 bits, weight 6) instead of a 2030-bit
// square-and-multiply.  The result is the CUBE of the reduced pairing; since 3 does not divide r it is one
// exactly when the pairing product is, which is all batch_check asks.  After the easy part elements are unitary
// (conjugate = inverse), so negative exponents are conjugations.
struct Frob { f12 m1[12], m2[12]; };   // (w^i)^p, (w^i)^(p^2)
static f12 f12_pow_words(const f12& a, const uint64_t* e, int words) {
    f12 acc = f12_one();
    for (int i = 64 * words - 1; i >= 0; i--) {
        acc = f12_mul(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = f12_mul(acc, a);
    }
    return acc;
}
static f12 frob_apply(const f12 m[12], const f12& x) {
    f12 r = f12_zero();
    for (int i = 0; i < 12; i++) {
        if (x.c[i].is_zero()) continue;
        for (int j = 0; j < 12; j++) r.c[j] = Q::add(r.c[j], Q::mul(x.c[i], m[i].c[j]));
    }
    return r;
}
static const Frob& FR() {
    static const Frob fr = [] {
        Frob t;
        f12 w = f12_zero();
        w.c[1] = Q::one();
        const f12 wp = f12_pow_words(w, Q::C().p, 6);
        t.m1[0] = f12_one();
        for (int i = 1; i < 12; i++) t.m1[i] = f12_mul(t.m1[i - 1], wp);
        for (int i = 0; i < 12; i++) t.m2[i] = frob_apply(t.m1, t.m1[i]);      // ((w^i)^p)^p
        return t;
    }();
    return fr;
}
static f12 pow_abs_x(const f12& a) {            // a^|x|, |x| = 0xd201000000010000
    static const uint64_t X = 0xd201000000010000ULL;
    f12 acc = a;
    for (int i = 62; i >= 0; i--) {
        acc = f12_mul(acc, acc);
        if ((X >> i) & 1) acc = f12_mul(acc, a);
    }
    return acc;
}
// f^(3 (p^12 - 1) / r)
static f12 final_exponentiation_cubed(const f12& f) {
    const Frob& fr = FR();
    f12 g = f12_mul(f12_conj(f), f12_inv(f));                   // f^(p^6 - 1)
    g = f12_mul(frob_apply(fr.m2, g), g);                        // ^(p^2 + 1): unitary from here on
    auto pow_x = [](const f12& a) { return f12_conj(pow_abs_x(a)); };               // a^x, x < 0
    auto pow_xm1 = [](const f12& a) { return f12_conj(f12_mul(pow_abs_x(a), a)); }; // a^(x - 1) = conj(a^(|x| + 1))
    const f12 a = pow_xm1(pow_xm1(g));                           // g^((x - 1)^2)
    const f12 b = f12_mul(pow_x(a), frob_apply(fr.m1, a));       // a^(x + p)
    const f12 c = f12_mul(f12_mul(pow_x(pow_x(b)), frob_apply(fr.m2, b)), f12_conj(b));   // b^(x^2 + p^2 - 1)
    return f12_mul(c, f12_mul(f12_mul(g, g), g));                // * g^3
}

static f12 final_exponentiation(const f12& f) {
    // (p^6 - 1): conjugate over inverse; then (p^6 + 1) / r by square-and-multiply
    static const uint64_t E[32] = {
    0x8739e1cdc0705d6aULL, 0x09a5256de0381a16ULL, 0x9cf0f70a61c791e2ULL, 0x3a09c4497903f76eULL,
    0x2d7271563890f133ULL, 0x224741b36fec7760ULL, 0x338259c22a12bd40ULL, 0x38ee1cd4778e0de7ULL,
    0xc3b5ef4b188a20b0ULL, 0x1d615d49e2764d7bULL, 0x816101ddd076117dULL, 0xf007c01e7ebe3afcULL,
    0x27d7bd90935021c3ULL, 0xc3b5e2f557c0b15fULL, 0x5e886c94c4f82384ULL, 0xee6a95db11e63f56ULL,
    0x2b822f514a9c4f6fULL, 0x12d6a874d21b73daULL, 0x1304275ef499dffbULL, 0x967878febcb95d1fULL,
    0x4744497f8b2f2922ULL, 0x85a2e707f0841855ULL, 0x9f0c50126c802eecULL, 0xfb46e197bd2fa489ULL,
    0x548ce0809bc5f61aULL, 0xcf56fb1573beaa8cULL, 0xad7375a3763bdf7cULL, 0xe0ec9031179bdeccULL,
    0x6579aea83c48c1daULL, 0xdbf85ae664cf5bb3ULL, 0x7b6f235c55ca7566ULL, 0x000028b314877503ULL};
    const f12 g = f12_mul(f12_conj(f), f12_inv(f));
    f12 out = f12_one();
    for (int i = 2029; i >= 0; i--) {
        out = f12_mul(out, out);
        if ((E[i >> 6] >> (i & 63)) & 1) out = f12_mul(out, g);
    }
    return out;
}

// prod_i e(P_i, Q_i) == 1 (one shared final exponentiation: multi_miller_loop(..).final_exp())
static bool pairing_product_is_one(const P1* ps, const P2* qs, size_t m) {
    f12 f = f12_one();
    for (size_t i = 0; i < m; i++) {
        fq x, y; bool inf;
        p1_to_affine(ps[i], &x, &y, &inf);
        f = f12_mul(f, miller_loop(qs[i], x, y, inf));
    }
    return f12_is_one(final_exponentiation_cubed(f));
}

// test hook: the Frobenius-based exponentiation equals the cube of the plain 2030-bit square-and-multiply
static bool final_exponentiations_agree(const P1& p, const P2& q) {
    fq x, y; bool inf;
    p1_to_affine(p, &x, &y, &inf);
    const f12 f = miller_loop(q, x, y, inf);
    const f12 slow = final_exponentiation(f);
    const f12 fast = final_exponentiation_cubed(f);
    const f12 cube = f12_mul(f12_mul(slow, slow), slow);
    for (int i = 0; i < 12; i++) if (!fq_eq(cube.c[i], fast.c[i])) return false;
    return !f12_is_one(slow);
}

// batch_check (src/commitment_scheme.rs:24-66) over `count` flattened opening proofs
static bool batch_check(const P2& beta_h, const fr* points, const P1* witness, const fr* evals, const P1* polys,
                        size_t count, drv::Transcript& tr) {
    P1 total_c = p1_inf(), total_w = p1_inf();
    const fr u = tr.challenge_scalar("batch");
    fr pw = F::one(), g_mult = F::zero();
    for (size_t i = 0; i < count; i++) {
        P1 c = polys[i];
        p1_add(c, p1_mul(witness[i], points[i]));          // c += w * point
        g_mult = F::add(g_mult, F::mul(pw, evals[i]));
        p1_add(total_c, p1_mul(c, pw));
        p1_add(total_w, p1_mul(witness[i], pw));
        pw = F::mul(pw, u);
    }
    p1_add(total_c, p1_neg(p1_mul(p1_generator(), g_mult)));
    const P1 ps[2] = {p1_neg(total_w), total_c};
    const P2 qs[2] = {beta_h, p2_generator()};
    return pairing_product_is_one(ps, qs, 2);
}

}  // namespace vf
}  // namespace zkp

using namespace zkp;
using zkp::drv::F;
using zkp::drv::fr;
using zkp::drv::fr_load;

extern "C" {

int zkp_g1_decompress(const uint8_t in[48], uint64_t out_xy[12]) {
    if (!in || !out_xy) return ZKP_ERR_INVALID;
    return vf::g1_decompress(in, out_xy) ? ZKP_OK : ZKP_ERR_INVALID;
}

/* Proof decoding (src/prover/proof.rs:36-66 derives Decode; "subgroup checks are done when the proof is deserialized",
 * proof.rs:77): 11 compressed G1 + 16 canonical little-endian scalars -> what zkp_verify takes. */
int zkp_proof_decode(const uint8_t bytes[1040], uint64_t commitments[132], uint64_t evaluations[64]) {
    if (!bytes || !commitments || !evaluations) return ZKP_ERR_INVALID;
    for (int j = 0; j < 11; j++)
        if (!vf::g1_decompress(bytes + 48 * j, commitments + 12 * j)) return ZKP_ERR_INVALID;
    // wire order of `Evaluations` (linearization_poly.rs:113-130) -> EVAL order of zkp_prover_prove
    static const int wire_order[16] = {0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 7, 8, 9, 15, 14};
    for (int j = 0; j < 16; j++) {
        fr v;
        memset(v.l, 0, 32);
        const uint8_t* b = bytes + 48 * 11 + 32 * j;
        for (int i = 0; i < 32; i++) v.l[i / 8] |= (uint64_t)b[i] << (8 * (i % 8));
        if (F::geq_p(v.l)) return ZKP_ERR_INVALID;          // non-canonical scalar
        drv::fr_store(evaluations + 4 * wire_order[j], F::to_mont(v));
    }
    return ZKP_OK;
}

int zkp_g2_generator_mul(const uint64_t scalar[4], uint64_t out[24]) {
    if (!scalar || !out) return ZKP_ERR_INVALID;
    vf::p2_store(vf::p2_mul(vf::p2_generator(), fr_load(scalar)), out);
    return ZKP_OK;
}

int zkp_g1_generator_mul(const uint64_t scalar[4], uint64_t out[12]) {
    if (!scalar || !out) return ZKP_ERR_INVALID;
    vf::fq x, y; bool inf;
    vf::p1_to_affine(vf::p1_mul(vf::p1_generator(), fr_load(scalar)), &x, &y, &inf);
    memcpy(out, x.l, 48); memcpy(out + 6, y.l, 48);
    return ZKP_OK;
}

/* self-check of the two final exponentiations on one pairing (tests/test_verifier_host.py) */
int zkp_pairing_selftest(const uint64_t g1[12], const uint64_t g2[24]) {
    if (!g1 || !g2) return ZKP_ERR_INVALID;
    const vf::P2 q = vf::p2_load(g2);
    if (!vf::p2_on_curve(q)) return ZKP_ERR_INVALID;
    return vf::final_exponentiations_agree(vf::p1_affine(g1), q) ? ZKP_OK : ZKP_ERR_VERIFY;
}

int zkp_pairing_check(const uint64_t* g1, const uint64_t* g2, size_t count) {
    if ((!g1 || !g2) && count) return ZKP_ERR_INVALID;
    std::vector<vf::P1> ps(count);
    std::vector<vf::P2> qs(count);
    for (size_t i = 0; i < count; i++) {
        ps[i] = vf::p1_affine(g1 + 12 * i);
        qs[i] = vf::p2_load(g2 + 24 * i);
        if (!vf::p2_on_curve(qs[i])) return ZKP_ERR_INVALID;
    }
    return vf::pairing_product_is_one(ps.data(), qs.data(), count) ? ZKP_OK : ZKP_ERR_VERIFY;
}

int zkp_kzg_batch_check(const uint64_t beta_h[24], const uint64_t* points, const uint64_t* witness_comms,
                        const uint64_t* evals, const uint64_t* poly_comms, size_t count, uint8_t transcript[203]) {
    if (!beta_h || !points || !witness_comms || !evals || !poly_comms || !transcript || count == 0 || count > 64)
        return ZKP_ERR_INVALID;
    const vf::P2 bh = vf::p2_load(beta_h);
    if (!vf::p2_on_curve(bh)) return ZKP_ERR_INVALID;
    std::vector<fr> pts(count), evs(count);
    std::vector<vf::P1> ws(count), cs(count);
    for (size_t i = 0; i < count; i++) {
        pts[i] = fr_load(points + 4 * i); evs[i] = fr_load(evals + 4 * i);
        ws[i] = vf::p1_affine(witness_comms + 12 * i); cs[i] = vf::p1_affine(poly_comms + 12 * i);
    }
    drv::Transcript tr;
    tr.load(transcript);
    const bool ok = vf::batch_check(bh, pts.data(), ws.data(), evs.data(), cs.data(), count, tr);
    tr.save(transcript);
    return ok ? ZKP_OK : ZKP_ERR_VERIFY;
}

int zkp_verify(const zkp_verifier_key* vk, const uint64_t beta_h[24], const uint8_t transcript[203],
               const uint64_t commitments[132], const uint64_t evaluations[64], const uint32_t* pi_idx,
               const uint64_t* pi_values, size_t pi_count) {
    if (!vk || !beta_h || !transcript || !commitments || !evaluations || vk->k < 1 || vk->k > 28 ||
        ((!pi_idx || !pi_values) && pi_count))
        return ZKP_ERR_INVALID;
    const uint64_t n = 1ull << vk->k;
    for (size_t i = 0; i < pi_count; i++) if (pi_idx[i] >= n) return ZKP_ERR_INVALID;
    const vf::P2 bh = vf::p2_load(beta_h);
    if (!vf::p2_on_curve(bh)) return ZKP_ERR_INVALID;
    drv::Transcript tr;
    tr.load(transcript);
    // Verifier::verify (src/verifier.rs:60-67): the public inputs enter the transcript first
    for (size_t i = 0; i < pi_count; i++) tr.append_scalar("pi", fr_load(pi_values + 4 * i));
    enum { A = 0, B, C, D, Z, TLO, TMID, THI, T4, WZ, WZW };
    enum { E_A = 0, E_B, E_C, E_D, E_AN, E_BN, E_DN, E_S1, E_S2, E_S3, E_QARITH, E_QC, E_QL, E_QR, E_PERM, E_R };
    auto cm = [&](int i) { return commitments + 12 * i; };
    fr e[16];
    for (int i = 0; i < 16; i++) e[i] = fr_load(evaluations + 4 * i);
    // Proof::verify (src/prover/proof.rs:70-383)
    static const char* const wl[4] = {"a_w", "b_w", "c_w", "d_w"};
    for (int j = 0; j < 4; j++) tr.append_commitment(wl[j], cm(A + j));
    const fr beta = tr.challenge_scalar("beta");
    tr.append_scalar("beta", beta);
    const fr gamma = tr.challenge_scalar("gamma");
    tr.append_commitment("z", cm(Z));
    fr ch[8];
    ch[0] = tr.challenge_scalar("alpha");
    ch[1] = beta; ch[2] = gamma;
    ch[3] = tr.challenge_scalar("range separation challenge");
    ch[4] = tr.challenge_scalar("logic separation challenge");
    ch[5] = tr.challenge_scalar("fixed base separation challenge");
    ch[6] = tr.challenge_scalar("variable base separation challenge");
    static const char* const tl[4] = {"t_low", "t_mid", "t_high", "t_4"};
    for (int j = 0; j < 4; j++) tr.append_commitment(tl[j], cm(TLO + j));
    const fr zc = tr.challenge_scalar("z_challenge");
    ch[7] = zc;
    const fr alpha = ch[0], one = F::one();
    // domain constants of VerificationKey (src/key.rs:203-214): n^-1, generator, generator^-1
    const fr n_fr = F::from_u64(n), n_inv = F::inv(n_fr);
    fr gen;
    {
        const fr_t g = fft_constant_host(vk->k, 0);
        memcpy(gen.l, g.l, 32);
    }
    const fr gen_inv = F::inv(gen);
    const fr z_n = F::pow(zc, n);
    const fr z_h = F::sub(z_n, one);
    if (z_h.is_zero() || F::sub(zc, one).is_zero()) return ZKP_ERR_VERIFY;
    const fr l1 = F::mul(z_h, F::inv(F::mul(n_fr, F::sub(zc, one))));
    // barycentric PI(z) over the non-zero public inputs (src/prover/proof.rs:541-591)
    fr pi_eval = F::zero();
    for (size_t i = 0; i < pi_count; i++) {
        const fr v = fr_load(pi_values + 4 * i);
        if (v.is_zero()) continue;
        const fr den = F::sub(F::mul(F::pow(gen_inv, pi_idx[i]), zc), one);
        if (den.is_zero()) return ZKP_ERR_VERIFY;
        pi_eval = F::add(pi_eval, F::mul(v, F::inv(den)));
    }
    pi_eval = F::mul(pi_eval, F::mul(z_h, n_inv));
    // compute_quotient_evaluation (src/prover/proof.rs:386-440)
    const fr a_ = F::add(e[E_R], pi_eval);
    const fr b0 = F::add(F::add(e[E_A], F::mul(beta, e[E_S1])), gamma);
    const fr b1 = F::add(F::add(e[E_B], F::mul(beta, e[E_S2])), gamma);
    const fr b2 = F::add(F::add(e[E_C], F::mul(beta, e[E_S3])), gamma);
    const fr b3 = F::mul(F::mul(F::add(e[E_D], gamma), e[E_PERM]), alpha);
    const fr b_ = F::mul(F::mul(b0, b1), F::mul(b2, b3));
    const fr c_ = F::mul(l1, F::sqr(alpha));
    const fr t_eval = F::mul(F::sub(F::sub(a_, b_), c_), F::inv(z_h));
    // compute_quotient_commitment (src/prover/proof.rs:442-455)
    vf::P1 t_comm = vf::p1_affine(cm(TLO));
    vf::p1_add(t_comm, vf::p1_mul(vf::p1_affine(cm(TMID)), z_n));
    vf::p1_add(t_comm, vf::p1_mul(vf::p1_affine(cm(THI)), F::sqr(z_n)));
    vf::p1_add(t_comm, vf::p1_mul(vf::p1_affine(cm(T4)), F::mul(F::sqr(z_n), z_n)));
    static const char* const el[15] = {"a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval",
                                       "d_next_eval", "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval",
                                       "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "perm_eval"};
    for (int j = 0; j < 15; j++) tr.append_scalar(el[j], e[j]);
    tr.append_scalar("t_eval", t_eval);
    tr.append_scalar("r_eval", e[E_R]);
    // compute_linearization_commitment: the prover's linearisation scalars on the committed polynomials
    fr sc[12];
    drv::linearization_scalars(n, ch, e, sc);
    static const int lin_vk[12] = {0, 1, 2, 3, 5, 4, 7, 8, 9, 10, -1, 14};   // q_m q_l q_r q_o q_4 q_c q_range q_logic q_fixed q_var z s_sigma_4
    vf::P1 r_comm = vf::p1_inf();
    for (int j = 0; j < 12; j++) {
        const vf::P1 pt = lin_vk[j] < 0 ? vf::p1_affine(cm(Z)) : vf::p1_affine(vk->commitments[lin_vk[j]]);
        vf::p1_add(r_comm, vf::p1_mul(pt, sc[j]));
    }
    // AggregateProof::flatten (src/commitment_scheme.rs:104-153)
    auto flatten = [&](const fr* evs, const vf::P1* pts, int cnt, fr* ev_out, vf::P1* cm_out) {
        const fr v = tr.challenge_scalar("v_challenge");
        fr pw = F::one(), acc = F::zero();
        vf::P1 c = vf::p1_inf();
        for (int i = 0; i < cnt; i++) {
            vf::p1_add(c, vf::p1_mul(pts[i], pw));
            acc = F::add(acc, F::mul(evs[i], pw));
            pw = F::mul(pw, v);
        }
        *ev_out = acc; *cm_out = c;
    };
    const fr ea[9] = {t_eval, e[E_R], e[E_A], e[E_B], e[E_C], e[E_D], e[E_S1], e[E_S2], e[E_S3]};
    const vf::P1 pa[9] = {t_comm, r_comm, vf::p1_affine(cm(A)), vf::p1_affine(cm(B)), vf::p1_affine(cm(C)),
                          vf::p1_affine(cm(D)), vf::p1_affine(vk->commitments[11]), vf::p1_affine(vk->commitments[12]),
                          vf::p1_affine(vk->commitments[13])};
    const fr eb[4] = {e[E_PERM], e[E_AN], e[E_BN], e[E_DN]};
    const vf::P1 pb[4] = {vf::p1_affine(cm(Z)), vf::p1_affine(cm(A)), vf::p1_affine(cm(B)), vf::p1_affine(cm(D))};
    fr evs[2];
    vf::P1 polys[2];
    flatten(ea, pa, 9, &evs[0], &polys[0]);
    flatten(eb, pb, 4, &evs[1], &polys[1]);
    tr.append_commitment("w_z", cm(WZ));
    tr.append_commitment("w_z_w", cm(WZW));
    const fr points[2] = {zc, F::mul(zc, gen)};
    const vf::P1 wit[2] = {vf::p1_affine(cm(WZ)), vf::p1_affine(cm(WZW))};
    return vf::batch_check(bh, points, wit, evs, polys, 2, tr) ? ZKP_OK : ZKP_ERR_VERIFY;
}

}  // extern "C"

// Modular inversion by an optimized binary GCD (Pornin, "Optimized Binary GCD for Modular Inversion",
// ePrint 2020/972) on 64-bit limbs, host code.  The prover inverts on the host at every synchronisation
// point of the transcript (one Fq inversion per commit group, Commitment::new; one Fr inversion in the
// permutation product), where a Fermat chain costs 570 (Fq) / 380 (Fr) dependent multiplications; this
// runs the 2 len(m) - 1 binary-GCD steps in ceil((2 len - 1) / 31) rounds of 31 steps on 64-bit
// approximations of (a, b), applying each round's 2 x 2 update matrix to the full-length values at once.
//
// inv_mod(y, m, out): out = y^-1 mod m for 0 < y < m, m odd; out = 0 for y = 0.  Plain integers (the
// caller handles the Montgomery factors).
#pragma once
#include <stdint.h>
#include <string.h>

namespace zkp {
namespace hostinv {

typedef unsigned __int128 u128;
typedef __int128 i128;

template <int N>
static inline int bit_length(const uint64_t* a) {
    for (int i = N - 1; i >= 0; i--)
        if (a[i]) return 64 * i + 64 - __builtin_clzll(a[i]);
    return 0;
}

// bits [pos, pos + 33) of a (pos >= 0)
template <int N>
static inline uint64_t bits33(const uint64_t* a, int pos) {
    const int w = pos >> 6, off = pos & 63;
    u128 v = a[w];
    if (w + 1 < N) v |= (u128)a[w + 1] << 64;
    return (uint64_t)(v >> off) & ((1ull << 33) - 1);
}

// r (N + 1 limbs, two's complement) = f x + g y for N-limb unsigned x, y and |f|, |g| <= 2^31
template <int N>
static inline void lin2(uint64_t* r, int64_t f, const uint64_t* x, int64_t g, const uint64_t* y) {
    i128 acc = 0;
    for (int i = 0; i < N; i++) {
        acc += (i128)f * (i128)(u128)x[i] + (i128)g * (i128)(u128)y[i];
        r[i] = (uint64_t)acc;
        acc >>= 64;   // arithmetic
    }
    r[N] = (uint64_t)acc;
}

// x (N + 1 limbs, two's complement) >>= 31 (arithmetic); returns the result's sign (1 = negative)
template <int N>
static inline int shr31(uint64_t* x) {
    for (int i = 0; i < N; i++) x[i] = (x[i] >> 31) | (x[i + 1] << 33);
    x[N] = (uint64_t)((int64_t)x[N] >> 31);
    return (int)(x[N] >> 63);
}

template <int N>
static inline void negate(uint64_t* x) {   // N + 1 limbs
    uint64_t c = 1;
    for (int i = 0; i <= N; i++) { const uint64_t v = ~x[i] + c; c = (c && v == 0) ? 1 : 0; x[i] = v; }
}

template <int N>
void inv_mod(const uint64_t* y, const uint64_t* m, uint64_t* out) {
    uint64_t a[N + 1], b[N + 1], u[N + 1], v[N + 1], ta[N + 1], tb[N + 1];
    memcpy(a, y, 8 * N); a[N] = 0;
    memcpy(b, m, 8 * N); b[N] = 0;
    memset(u, 0, sizeof u); u[0] = 1;
    memset(v, 0, sizeof v);
    // -m^-1 mod 2^31 (Newton on the low limb)
    uint64_t mi = 1;
    for (int i = 0; i < 6; i++) mi *= 2 - m[0] * mi;
    const uint64_t m_neg_inv31 = (0 - mi) & 0x7fffffffull;
    const int rounds = (2 * bit_length<N>(m) - 1 + 30) / 31;
    for (int round = 0; round < rounds; round++) {
        // 64-bit approximations: the low 31 bits and the 33 bits below the common top
        const int la = bit_length<N>(a), lb = bit_length<N>(b);
        const int n = la > lb ? la : lb;
        uint64_t xa, xb;
        if (n <= 64) { xa = a[0]; xb = b[0]; }
        else {
            xa = (a[0] & 0x7fffffffull) | (bits33<N>(a, n - 33) << 31);
            xb = (b[0] & 0x7fffffffull) | (bits33<N>(b, n - 33) << 31);
        }
        int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
        for (int j = 0; j < 31; j++) {
            if (xa & 1) {
                if (xa < xb) {
                    uint64_t t = xa; xa = xb; xb = t;
                    int64_t s = f0; f0 = f1; f1 = s;
                    s = g0; g0 = g1; g1 = s;
                }
                xa -= xb; f0 -= f1; g0 -= g1;
            }
            xa >>= 1; f1 <<= 1; g1 <<= 1;
        }
        // (a, b) <- (f0 a + g0 b, f1 a + g1 b) / 2^31, made non-negative
        lin2<N>(ta, f0, a, g0, b);
        lin2<N>(tb, f1, a, g1, b);
        if (shr31<N>(ta)) { negate<N>(ta); f0 = -f0; g0 = -g0; }
        if (shr31<N>(tb)) { negate<N>(tb); f1 = -f1; g1 = -g1; }
        // (u, v) <- (f0 u + g0 v, f1 u + g1 v) / 2^31 mod m
        uint64_t tu[N + 1], tv[N + 1];
        lin2<N>(tu, f0, u, g0, v);
        lin2<N>(tv, f1, u, g1, v);
        uint64_t* tt[2] = {tu, tv};
        for (int k = 0; k < 2; k++) {
            uint64_t* t = tt[k];
            // add the multiple of m that clears the low 31 bits
            const uint64_t q = (t[0] * m_neg_inv31) & 0x7fffffffull;
            u128 c = 0;
            for (int i = 0; i < N; i++) { c += (u128)q * m[i] + t[i]; t[i] = (uint64_t)c; c >>= 64; }
            t[N] += (uint64_t)c;
            shr31<N>(t);
            // now -m < t < 2 m: bring into [0, m)
            if (t[N] >> 63) {
                u128 cc = 0;
                for (int i = 0; i < N; i++) { cc += (u128)t[i] + m[i]; t[i] = (uint64_t)cc; cc >>= 64; }
                t[N] += (uint64_t)cc;
            } else {
                bool ge = t[N] != 0;
                if (!ge) {
                    ge = true;
                    for (int i = N - 1; i >= 0; i--) { if (t[i] > m[i]) break; if (t[i] < m[i]) { ge = false; break; } }
                }
                if (ge) {
                    uint64_t br = 0;
                    for (int i = 0; i < N; i++) { const u128 d = (u128)t[i] - m[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
                    t[N] -= br;
                }
            }
        }
        memcpy(a, ta, sizeof a); memcpy(b, tb, sizeof b);
        memcpy(u, tu, sizeof u); memcpy(v, tv, sizeof v);
    }
    // gcd in b (1 when y is invertible), inverse in v; y = 0 leaves b = m and v = 0
    memcpy(out, v, 8 * N);
}

}  // namespace hostinv
}  // namespace zkp

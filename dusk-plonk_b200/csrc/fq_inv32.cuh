// Divergence-free binary-GCD modular inversion on 32-bit limbs (Pornin, ePrint 2020/972, with k = 16:
// 32-bit approximations, 15 steps per round) -- EXPERIMENT / groundwork for round 2, not used by the product.
// A batched-affine bucket accumulation needs an Fq inversion far cheaper than the 570 dependent
// multiplications of a Fermat chain; this routine costs ceil((2 len(m) - 1) / 15) = 51 rounds of 15 cheap
// steps on one-word approximations plus four N-limb multiply-accumulates by 16-bit factors per round, with
// the same instruction stream for every input (selects, no data-dependent branches).
// The same source compiles for the host (bench/fq_inv32_host_test.cpp checks it against big integers).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FQI_HD __host__ __device__ __forceinline__
#else
#define FQI_HD inline
#endif

namespace fqinv {

template <int N>
FQI_HD int bit_length(const uint32_t* a) {
    int len = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
#if defined(__CUDA_ARCH__)
        const int l = 32 - __clz((int)a[i]);
#else
        const int l = a[i] ? 32 - __builtin_clz(a[i]) : 0;
#endif
        len = a[i] ? 32 * i + l : len;
    }
    return len;
}

// bits [pos, pos + 17) of a
template <int N>
FQI_HD uint32_t bits17(const uint32_t* a, int pos) {
    const int w = pos >> 5, off = pos & 31;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {   // select limbs w and w + 1 without dynamic indexing
        v |= (i == w) ? (uint64_t)a[i] : 0;
        v |= (i == w + 1) ? (uint64_t)a[i] << 32 : 0;
    }
    return (uint32_t)(v >> off) & 0x1ffffu;
}

// r (N + 1 limbs, two's complement) = f x + g y, |f|, |g| <= 2^15
template <int N>
FQI_HD void lin2(uint32_t* r, int32_t f, const uint32_t* x, int32_t g, const uint32_t* y) {
    int64_t acc = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        acc += (int64_t)f * (int64_t)x[i] + (int64_t)g * (int64_t)y[i];
        r[i] = (uint32_t)acc;
        acc >>= 32;
    }
    r[N] = (uint32_t)acc;
}

// x (N + 1 limbs) >>= 15, arithmetic; returns 1 if the result is negative
template <int N>
FQI_HD uint32_t shr15(uint32_t* x) {
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = (x[i] >> 15) | (x[i + 1] << 17);
    x[N] = (uint32_t)((int32_t)x[N] >> 15);
    return x[N] >> 31;
}

// x <- neg ? -x : x   (N + 1 limbs)
template <int N>
FQI_HD void cond_negate(uint32_t* x, uint32_t neg) {
    const uint32_t mask = 0u - neg;
    uint64_t c = neg;
#pragma unroll
    for (int i = 0; i <= N; i++) {
        c += (uint64_t)(x[i] ^ mask);
        x[i] = (uint32_t)c;
        c >>= 32;
    }
}

// t (N + 1 limbs, |t| < 2^15 m) <- t / 2^15 mod m, in [0, m): add the multiple of m that clears the low 15
// bits, shift, then one conditional correction (the quotient lies in (-m, 2m))
template <int N>
FQI_HD void div15_mod(uint32_t* t, const uint32_t* m, uint32_t m_neg_inv15) {
    const uint32_t q = (t[0] * m_neg_inv15) & 0x7fffu;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) { c += (uint64_t)q * m[i] + t[i]; t[i] = (uint32_t)c; c >>= 32; }
    t[N] += (uint32_t)c;
    shr15<N>(t);
    const uint32_t neg = t[N] >> 31;
    uint32_t s[N + 1];
    int64_t d = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {   // s = t + (neg ? m : -m)
        d += (int64_t)t[i] + (neg ? (int64_t)m[i] : -(int64_t)m[i]);
        s[i] = (uint32_t)d;
        d >>= 32;
    }
    s[N] = t[N] + (uint32_t)d;
    const bool take = neg || !(s[N] >> 31);   // negative: t + m; else t - m if that is still >= 0
#pragma unroll
    for (int i = 0; i <= N; i++) t[i] = take ? s[i] : t[i];
}

// out = y^-1 mod m (plain integers, N 32-bit limbs), m odd, 0 <= y < m; y = 0 gives 0.
// m_neg_inv15 = -m^-1 mod 2^15.
template <int N>
FQI_HD void inv_mod(const uint32_t* y, const uint32_t* m, uint32_t m_neg_inv15, int rounds, uint32_t* out) {
    uint32_t a[N + 1], b[N + 1], u[N + 1], v[N + 1];
#pragma unroll
    for (int i = 0; i < N; i++) { a[i] = y[i]; b[i] = m[i]; u[i] = 0; v[i] = 0; }
    a[N] = b[N] = u[N] = v[N] = 0;
    u[0] = 1;
    for (int round = 0; round < rounds; round++) {
        const int la = bit_length<N>(a), lb = bit_length<N>(b);
        const int n = la > lb ? la : lb;
        const int pos = n > 32 ? n - 17 : 15;   // n <= 32: bits [15, 32) of the low limb
        uint32_t xa = (a[0] & 0x7fffu) | (bits17<N>(a, pos) << 15);
        uint32_t xb = (b[0] & 0x7fffu) | (bits17<N>(b, pos) << 15);
        int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll
        for (int j = 0; j < 15; j++) {
            const uint32_t odd = xa & 1u;
            const bool sw = odd && xa < xb;
            const uint32_t ta = sw ? xb : xa, tb = sw ? xa : xb;
            const int32_t tf0 = sw ? f1 : f0, tf1 = sw ? f0 : f1, tg0 = sw ? g1 : g0, tg1 = sw ? g0 : g1;
            xa = (ta - (odd ? tb : 0u)) >> 1;
            xb = tb;
            f0 = tf0 - (odd ? tf1 : 0);
            g0 = tg0 - (odd ? tg1 : 0);
            f1 = tf1 << 1;
            g1 = tg1 << 1;
        }
        uint32_t ta[N + 1], tb[N + 1], tu[N + 1], tv[N + 1];
        lin2<N>(ta, f0, a, g0, b);
        lin2<N>(tb, f1, a, g1, b);
        const uint32_t na = shr15<N>(ta), nb = shr15<N>(tb);
        cond_negate<N>(ta, na);
        cond_negate<N>(tb, nb);
        f0 = na ? -f0 : f0; g0 = na ? -g0 : g0;
        f1 = nb ? -f1 : f1; g1 = nb ? -g1 : g1;
        lin2<N>(tu, f0, u, g0, v);
        lin2<N>(tv, f1, u, g1, v);
        div15_mod<N>(tu, m, m_neg_inv15);
        div15_mod<N>(tv, m, m_neg_inv15);
#pragma unroll
        for (int i = 0; i <= N; i++) { a[i] = ta[i]; b[i] = tb[i]; u[i] = tu[i]; v[i] = tv[i]; }
    }
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = v[i];
}

}  // namespace fqinv

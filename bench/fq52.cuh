// Fq (BLS12-381 base field) Montgomery multiplication on the FP64 pipe -- EXPERIMENT for round 2, not used
// by the product.  Radix 2^52, 8 limbs, R = 2^416.  A 52 x 52 -> 104-bit limb product is split exactly into
// its high and low 52 bits by two round-toward-zero FMAs (Emmart's floating-point big-integer scheme):
//     2^104 + hi52 * 2^52 = fma_rz(a, b, 2^104)                     (the ulp of [2^104, 2^105) is 2^52)
//     2^52 + lo52         = fma_rz(a, b, (2^104 + 2^52) - that)     (the addend cancels hi52 * 2^52 exactly)
// and both halves are accumulated as 64-bit integers straight from the bit patterns (the exponent fields
// are constant, so adding bit patterns adds mantissas; the constants are subtracted once per column).
// Compiles for the host too (fesetround(FE_TOWARDZERO) + fma): bench/fq52_host_test.cpp checks exactness
// against the 32-bit-limb Montgomery multiplication of the product (arith.cuh).
#pragma once
#include <stdint.h>
#include <string.h>
#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

#if defined(__CUDACC__)
#define FQ52_HD __host__ __device__ __forceinline__
#else
#define FQ52_HD inline
#endif

namespace fq52 {

static constexpr int N = 8;
static constexpr uint64_t MASK = (1ull << 52) - 1;
static constexpr uint64_t BITS_C1 = 0x4670000000000000ull;   // bit pattern of 2^104
static constexpr uint64_t BITS_2P52 = 0x4330000000000000ull; // bit pattern of 2^52

FQ52_HD double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    return fma(a, b, c);   // the caller has set FE_TOWARDZERO
#endif
}
FQ52_HD uint64_t bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
FQ52_HD double from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
// exact double of an integer below 2^52 without an int -> float conversion instruction
FQ52_HD double to_double(uint64_t x) { return from_bits(x | BITS_2P52) - 4503599627370496.0; }

// p in 52-bit limbs, and -p^-1 mod 2^52
FQ52_HD uint64_t P52(int i) {
    constexpr uint64_t t[8] = {0xeffffffffaaabull, 0xfeb153ffffb9full, 0x6b0f6241eabffull, 0x12bf6730d2a0full,
                               0x764774b84f385ull, 0x1ba7b6434bacdull, 0x1ea397fe69a4bull, 0x1a011ull};
    return t[i];
}
static constexpr uint64_t PINV52 = 0x3fffcfffcfffdull;

struct el { double l[N]; };   // limbs as exact doubles, value < p

// col[k] += lo52(a b), col[k + 1] += hi52(a b), as raw bit patterns (constants removed by the caller)
FQ52_HD void mul_acc(uint64_t* col, int k, double a, double b) {
    const double c1 = 20282409603651670423947251286016.0;                       // 2^104
    const double c2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52
    const double hi = fma_rz(a, b, c1);
    const double lo = fma_rz(a, b, c2 - hi);
    col[k + 1] += bits(hi);
    col[k] += bits(lo);
}

// Montgomery product a b R^-1 mod p, R = 2^416; inputs and output fully reduced (< p).
FQ52_HD el mul(const el& a, const el& b) {
    uint64_t col[2 * N + 1];
#pragma unroll
    for (int k = 0; k <= 2 * N; k++) col[k] = 0;
    // schoolbook product: column k receives lo parts of i + j = k and hi parts of i + j = k - 1
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) mul_acc(col, i + j, a.l[i], b.l[j]);
    // remove the bit-pattern constants: column k got min(k, 2N-2-k)+1 lo terms (k <= 2N-2) and as many hi
    // terms as column k-1 has lo terms
#pragma unroll
    for (int k = 0; k <= 2 * N; k++) {
        const int nlo = k <= 2 * N - 2 ? (k < N ? k + 1 : 2 * N - 1 - k) : 0;
        const int nhi = (k >= 1 && k - 1 <= 2 * N - 2) ? (k - 1 < N ? k : 2 * N - k) : 0;
        col[k] -= (uint64_t)nlo * BITS_2P52 + (uint64_t)nhi * BITS_C1;
    }
    // reduction: one limb at a time
    const double pinv = to_double(PINV52);
#pragma unroll
    for (int i = 0; i < N; i++) {
        col[i + 1] += col[i] >> 52;
        const uint64_t c = col[i] & MASK;
        // q = lo52(c * pinv)
        const double c1 = 20282409603651670423947251286016.0;
        const double c2 = 20282409603651670423947251286016.0 + 4503599627370496.0;
        const double cd = to_double(c);
        const double qh = fma_rz(cd, pinv, c1);
        const double ql = fma_rz(cd, pinv, c2 - qh);
        const double q = ql - 4503599627370496.0;
        uint64_t t[N + 1];
#pragma unroll
        for (int j = 0; j <= N; j++) t[j] = 0;
#pragma unroll
        for (int j = 0; j < N; j++) mul_acc(t, j, q, to_double(P52(j)));
        // t[0] has 1 lo, t[j] 1 lo + 1 hi, t[N] 1 hi
        t[0] -= BITS_2P52;
#pragma unroll
        for (int j = 1; j < N; j++) t[j] -= BITS_2P52 + BITS_C1;
        t[N] -= BITS_C1;
        // c + lo52(q p_0) = 0 mod 2^52: only its carry survives
        col[i + 1] += (c + t[0]) >> 52;
#pragma unroll
        for (int j = 1; j <= N; j++) col[i + j] += t[j];
    }
    // carry-normalise the upper half and subtract p once if needed
    uint64_t r[N];
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
        const uint64_t v = col[N + k] + carry;
        r[k] = v & MASK;
        carry = v >> 52;
    }
    // value = r (+ carry * 2^416, impossible for reduced inputs: a b / R + p < 2p < 2^416)
    uint64_t s[N];
    uint64_t borrow = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
        const uint64_t d = r[k] - P52(k) - borrow;
        s[k] = d & MASK;
        borrow = (d >> 63) & 1;
    }
    el out;
#pragma unroll
    for (int k = 0; k < N; k++) out.l[k] = to_double(borrow ? r[k] : s[k]);
    return out;
}

}  // namespace fq52

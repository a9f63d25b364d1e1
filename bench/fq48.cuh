// Fq Montgomery multiplication entirely on the FP64 pipe -- EXPERIMENT for round 2, not used by the product.
// Radix 2^48, 8 limbs, R = 2^384 (the same R as the 32-bit-limb arithmetic of arith.cuh, so results are
// directly comparable).  Limbs are exact integers held in doubles.  No integer instruction touches the
// multiply-accumulate path: high halves of the limb products accumulate inside the FMA itself,
//     H <- fma_rz(a, b, H),  H in [2^100, 2^101)  (ulp 2^48: the FMA adds floor(a b / 2^48) 2^48 exactly),
// the low half of each product is recovered exactly by a second FMA against the high half just added,
//     lo = fma_rz(a, b, H_old - H_new)            (0 <= lo < 2^48),
// and sums of up to 16 low halves (< 2^52) and of 16 high halves (< 2^52 after scaling) stay exact in the
// 53-bit mantissa.  4 FP64 operations per limb product, CIOS order over an 8-column window.
// Purpose: an Fq multiplication that leaves the IMAD pipe free, to run side by side with the integer one.
#pragma once
#include <stdint.h>
#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

#if defined(__CUDACC__)
#define FQ48_HD __host__ __device__ __forceinline__
#else
#define FQ48_HD inline
#endif

namespace fq48 {

static constexpr int N = 8;

FQ48_HD double fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    return fma(a, b, c);    // the caller has set FE_TOWARDZERO
#endif
}
FQ48_HD double add_rz(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rz(a, b);
#else
    return a + b;           // FE_TOWARDZERO
#endif
}

static constexpr double C100 = 1267650600228229401496703205376.0;   // 2^100
static constexpr double TWO48 = 281474976710656.0;                  // 2^48
static constexpr double INV48 = 1.0 / 281474976710656.0;            // 2^-48
static constexpr double PINV = (double)0xfffcfffcfffdull;           // -p^-1 mod 2^48

FQ48_HD double P48(int i) {
    constexpr double t[8] = {(double)0xffffffffaaabull, (double)0xb153ffffb9feull, (double)0xf6241eabfffeull,
                             (double)0x6730d2a0f6b0ull, (double)0x4b84f38512bfull, (double)0x434bacd76477ull,
                             (double)0xe69a4b1ba7b6ull, (double)0x1a0111ea397full};
    return t[i];
}

struct el { double l[N]; };   // 48-bit limbs as exact doubles, value < p

// column j of the window: L[j] += lo48(x y), H[j] += hi48(x y) 2^48
FQ48_HD void mac(double& H, double& L, double x, double y) {
    const double old = H;
    H = fma_rz(x, y, old);
    L += fma_rz(x, y, old - H);
}

// a b 2^-384 mod p for a, b < p
FQ48_HD el mul(const el& a, const el& b) {
    double H[N], L[N];
#pragma unroll
    for (int j = 0; j < N; j++) { H[j] = C100; L[j] = 0.0; }
    double carry = 0.0;   // integer carried into window column 0
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) mac(H[j], L[j], a.l[j], b.l[i]);
        // column 0: v0 = low + c 2^48
        const double v0 = L[0] + carry;
        const double c48 = add_rz(v0, C100) - C100;
        const double low = v0 - c48;
        // q = low * (-p^-1) mod 2^48
        const double h = fma_rz(low, PINV, C100);
        const double q = fma_rz(low, PINV, C100 - h);
#pragma unroll
        for (int j = 0; j < N; j++) mac(H[j], L[j], q, P48(j));
        // column 0 is now a multiple of 2^48: it and the high halves parked there move up one column
        carry = ((L[0] + carry) + (H[0] - C100)) * INV48;
#pragma unroll
        for (int j = 0; j + 1 < N; j++) { H[j] = H[j + 1]; L[j] = L[j + 1]; }
        H[N - 1] = C100;
        L[N - 1] = 0.0;
    }
    // carry-normalise the remaining window (the upper half of the product)
    el r;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double v = L[j] + carry;
        const double c48 = add_rz(v, C100) - C100;
        r.l[j] = v - c48;
        carry = (c48 + (H[j] - C100)) * INV48;
    }
    // r < 2p: subtract p once if r >= p
    el s;
    double borrow = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        double d = r.l[j] - P48(j) - borrow;
        borrow = d < 0.0 ? 1.0 : 0.0;
        s.l[j] = d + borrow * TWO48;
    }
    const bool keep = borrow != 0.0;   // r < p
#pragma unroll
    for (int j = 0; j < N; j++) r.l[j] = keep ? r.l[j] : s.l[j];
    return r;
}

}  // namespace fq48

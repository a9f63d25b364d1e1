// Multi-limb Montgomery arithmetic for BLS12-381 Fr (8 x u32) and Fq (12 x u32) on the
// sm_100a integer pipe.
//
// The device path is inline PTX: every 32x32->64 product is a mad.lo.cc / madc.hi.cc pair
// on an aligned (even, odd) accumulator pair so that ptxas can fuse it into a single
// IMAD.WIDE.U32 with the carry kept in a predicate; the two interleaved accumulators
// ("even" columns and "odd" columns) remove the carry ripple between neighbouring
// products.  The same source also compiles for the host, where the PTX carry flag is
// emulated by a thread-local variable: the algorithm (not just its result) is unit-tested
// on the CPU (tests/test_host_arith.py) and reused by the C++ runtime for table set-up.
//
// Replaces: the Fr / Fq arithmetic of the reference's absent `bls-12-381` crate
// (reference use: `BlsScalar` src/lib.rs:81; Montgomery layout pinned by src/lib.rs:583-588).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZKP_HD __host__ __device__ __forceinline__
#define ZKP_D __device__ __forceinline__
#else
#define ZKP_HD inline
#define ZKP_D inline
#endif

namespace zkp {

// ------------------------------------------------------------------ carry primitives
#if defined(__CUDA_ARCH__)
ZKP_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZKP_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKP_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKP_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKP_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKP_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZKP_D uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// Host emulation of the PTX condition-code register (one per thread).
inline uint32_t& cc_flag() { static thread_local uint32_t cc = 0; return cc; }
inline uint32_t emu_add(uint32_t a, uint32_t b, uint32_t cin, bool set) {
    uint64_t t = (uint64_t)a + b + cin; if (set) cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t emu_sub(uint32_t a, uint32_t b, uint32_t bin, bool set) {
    uint64_t t = (uint64_t)a - b - bin; if (set) cc_flag() = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu_add(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu_add(a, b, cc_flag(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return emu_add(a, b, cc_flag(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, cc_flag(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return emu_sub(a, b, cc_flag(), false); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, 0, true); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, cc_flag(), true); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, 0, true); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, cc_flag(), true); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_hi(a, b), c, cc_flag(), false); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return emu_add(mul_lo(a, b), c, cc_flag(), false); }
#endif

// ------------------------------------------------------------------ field parameters
// Limbs are returned by constexpr functions over function-local tables so that, after
// full unrolling, every modulus limb is an immediate operand in device code.
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;  // -r^-1 mod 2^32
    // r = ... ffffffff 00000001: the two low limbs are 1 and 2^32 - 1 and INV is -1, so the
    // Montgomery reduction needs no multiplier for them (m * 1 = m, m * (2^32 - 1) = (m << 32) - m,
    // m = -t0): 5 of the 17 multiplications of every reduction row become additions.
    static constexpr bool LOW_LIMBS_SPECIAL = true;
    static constexpr ZKP_HD uint32_t p(int i) {
        constexpr uint32_t t[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return t[i];
    }
    static constexpr ZKP_HD uint32_t one(int i) {  // R mod r
        constexpr uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return t[i];
    }
    static constexpr ZKP_HD uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return t[i];
    }
};

struct FqParams {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;  // -p^-1 mod 2^32
    static constexpr bool LOW_LIMBS_SPECIAL = false;
    static constexpr ZKP_HD uint32_t p(int i) {
        constexpr uint32_t t[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu,
                                    0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u,
                                    0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return t[i];
    }
    static constexpr ZKP_HD uint32_t one(int i) {  // R mod p
        constexpr uint32_t t[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu,
                                    0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u,
                                    0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return t[i];
    }
    static constexpr ZKP_HD uint32_t r2(int i) {  // R^2 mod p
        constexpr uint32_t t[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u,
                                    0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u,
                                    0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return t[i];
    }
};

// ------------------------------------------------------------------ Montgomery field
template <class P>
struct alignas(16) Mont {
    static constexpr int N = P::N;
    uint32_t l[N];

    static ZKP_HD Mont zero() { Mont r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
    static ZKP_HD Mont one() { Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r; }
    static ZKP_HD Mont r2() { Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r; }
    static ZKP_HD Mont modulus() { Mont r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::p(i);
        return r; }

    ZKP_HD bool is_zero() const { uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < N; i++) x |= l[i];
        return x == 0; }
    ZKP_HD bool operator==(const Mont& o) const { uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < N; i++) x |= l[i] ^ o.l[i];
        return x == 0; }
    ZKP_HD bool operator!=(const Mont& o) const { return !(*this == o); }
};

// r = (a >= p) ? a - p : a        (a < 2p)
template <class P>
ZKP_HD void mont_final_sub(uint32_t* a) {
    constexpr int N = P::N;
    uint32_t t[N];
    t[0] = sub_cc(a[0], P::p(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = subc_cc(a[i], P::p(i));
    uint32_t borrow = subc(0, 0);  // 0xffffffff if a < p
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = borrow ? a[i] : t[i];
}

template <class P>
ZKP_HD Mont<P> operator+(const Mont<P>& a, const Mont<P>& b) {
    constexpr int N = P::N;
    Mont<P> r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);  // p < 2^(32N-1): no carry out
    mont_final_sub<P>(r.l);
    return r;
}

template <class P>
ZKP_HD Mont<P> operator-(const Mont<P>& a, const Mont<P>& b) {
    constexpr int N = P::N;
    Mont<P> r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
    uint32_t mask = subc(0, 0);  // all ones when a < b
    r.l[0] = add_cc(r.l[0], P::p(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], P::p(i) & mask);
    r.l[N - 1] = addc(r.l[N - 1], P::p(N - 1) & mask);
    return r;
}

template <class P>
ZKP_HD Mont<P> neg(const Mont<P>& a) {
    constexpr int N = P::N;
    Mont<P> r;
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < N; i++) nz |= a.l[i];
    uint32_t mask = nz ? 0xffffffffu : 0u;
    r.l[0] = sub_cc(P::p(0) & mask, a.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(P::p(i) & mask, a.l[i]);
    r.l[N - 1] = subc(P::p(N - 1) & mask, a.l[N - 1]);
    return r;
}

template <class P>
ZKP_HD Mont<P> dbl(const Mont<P>& a) { return a + a; }

// acc[j], acc[j+1] += x(j) * m  for j = 0, 2, ..  (one carry chain over the aligned pairs)
// OFF selects the even (0) or odd (1) limbs of the modulus.
template <class P, int OFF>
ZKP_HD void cmad_mod(uint32_t* acc, uint32_t m) {
    constexpr int N = P::N;
    if (P::LOW_LIMBS_SPECIAL) {
        static_assert(!P::LOW_LIMBS_SPECIAL || (P::p(0) == 1u && P::p(1) == 0xffffffffu), "limb pattern");
        if (OFF == 0) {           // + m * 1
            acc[0] = add_cc(acc[0], m);
            acc[1] = addc_cc(acc[1], 0);
        } else {                  // + m * (2^32 - 1) = (m - [m != 0]) * 2^32 + (2^32 - m) mod 2^32
            const uint32_t lo = 0u - m;
            const uint32_t hi = m - (m != 0u ? 1u : 0u);
            acc[0] = add_cc(acc[0], lo);
            acc[1] = addc_cc(acc[1], hi);
        }
    } else {
        acc[0] = mad_lo_cc(P::p(OFF), m, acc[0]);
        acc[1] = madc_hi_cc(P::p(OFF), m, acc[1]);
    }
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = madc_lo_cc(P::p(j + OFF), m, acc[j]);
        acc[j + 1] = madc_hi_cc(P::p(j + OFF), m, acc[j + 1]);
    }
}

// One row of the interleaved CIOS: T <- (T >> 32 on the previous row) + a * bi, then
// T <- T + m * p with m chosen so the low limb vanishes.  T = even + odd * 2^32.
template <class P, bool FIRST>
ZKP_HD void mad_row_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
    if (FIRST) {
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            odd[j] = mul_lo(a[j + 1], bi);
            odd[j + 1] = mul_hi(a[j + 1], bi);
        }
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            even[j] = mul_lo(a[j], bi);
            even[j + 1] = mul_hi(a[j], bi);
        }
    } else {
        // fold the limb that the implicit right shift moves to position 0
        even[0] = add_cc(even[0], odd[1]);
        // odd <- (odd >> 64) + odd-limb products, carry continues from the add above
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            odd[j] = madc_lo_cc(a[j + 1], bi, odd[j + 2]);
            odd[j + 1] = madc_hi_cc(a[j + 1], bi, odd[j + 3]);
        }
        odd[N - 2] = madc_lo_cc(a[N - 1], bi, 0);
        odd[N - 1] = madc_hi(a[N - 1], bi, 0);
        // even += even-limb products; its carry-out lands on the top odd limb
        even[0] = mad_lo_cc(a[0], bi, even[0]);
        even[1] = madc_hi_cc(a[0], bi, even[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            even[j] = madc_lo_cc(a[j], bi, even[j]);
            even[j + 1] = madc_hi_cc(a[j], bi, even[j + 1]);
        }
        odd[N - 1] = addc(odd[N - 1], 0);
    }
    const uint32_t m = P::LOW_LIMBS_SPECIAL ? 0u - even[0] : even[0] * P::INV;
    cmad_mod<P, 1>(odd, m);
    cmad_mod<P, 0>(even, m);
    odd[N - 1] = addc(odd[N - 1], 0);
}

template <class P>
ZKP_HD Mont<P> operator*(const Mont<P>& a, const Mont<P>& b) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count required");
    uint32_t even[N], odd[N];
    mad_row_redc<P, true>(even, odd, a.l, b.l[0]);
    mad_row_redc<P, false>(odd, even, a.l, b.l[1]);
#pragma unroll
    for (int i = 2; i < N; i += 2) {
        mad_row_redc<P, false>(even, odd, a.l, b.l[i]);
        mad_row_redc<P, false>(odd, even, a.l, b.l[i + 1]);
    }
    // The last row ran with the roles swapped: T = odd + even * 2^32 and odd[0] == 0,
    // so the result is (odd >> 32) + even.
    Mont<P> r;
    r.l[0] = add_cc(odd[1], even[0]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(odd[k + 1], even[k]);
    r.l[N - 1] = addc(even[N - 1], 0);
    mont_final_sub<P>(r.l);
    return r;
}

template <class P>
ZKP_HD Mont<P> sqr(const Mont<P>& a) { return a * a; }

// Montgomery -> canonical (multiply by 1) and back (multiply by R^2)
template <class P>
ZKP_HD Mont<P> from_mont(const Mont<P>& a) {
    Mont<P> o = Mont<P>::zero(); o.l[0] = 1; return a * o;
}
template <class P>
ZKP_HD Mont<P> to_mont(const Mont<P>& a) { return a * Mont<P>::r2(); }

template <class P>
ZKP_HD Mont<P> from_u64(uint64_t v) {
    Mont<P> o = Mont<P>::zero(); o.l[0] = (uint32_t)v; o.l[1] = (uint32_t)(v >> 32);
    return to_mont(o);
}

template <class P>
ZKP_HD Mont<P> pow_u64(Mont<P> base, uint64_t e) {
    Mont<P> acc = Mont<P>::one();
    while (e) {
        if (e & 1) acc = acc * base;
        base = sqr(base);
        e >>= 1;
    }
    return acc;
}

// a^(p-2) (Fermat); a = 0 maps to 0.
template <class P>
ZKP_HD Mont<P> inverse(const Mont<P>& a) {
    constexpr int N = P::N;
    uint32_t e[N];  // p - 2
    uint32_t borrow = 2;
#pragma unroll
    for (int j = 0; j < N; j++) {
        uint32_t pj = P::p(j);
        e[j] = pj - borrow;
        borrow = (pj < borrow) ? 1u : 0u;
    }
    Mont<P> acc = Mont<P>::one();
    for (int i = 32 * N - 1; i >= 0; i--) {
        acc = sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * a;
    }
    return acc;
}

using fr_t = Mont<FrParams>;
using fq_t = Mont<FqParams>;

}  // namespace zkp

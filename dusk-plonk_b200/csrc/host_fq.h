// Host-side Fq (6 x u64 CIOS Montgomery): the inversions at the MSM's synchronisation points, the combination
// of per-GPU partial commitments, and the verifier's group / pairing arithmetic (verifier.cu).
#pragma once
#include <stdint.h>
#include <string.h>

#include "arith.cuh"
#include "host_inv.h"

namespace zkp {
namespace hostfq {
typedef unsigned __int128 u128;
static const uint64_t P[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                              0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t INV = 0x89f3fffcfffcfffdULL;
static const uint64_t ONE[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                                0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
struct fq { uint64_t l[6]; };
static inline bool geq_p(const uint64_t* a) {
    for (int i = 5; i >= 0; i--) { if (a[i] > P[i]) return true; if (a[i] < P[i]) return false; }
    return true;
}
static inline fq mul(const fq& a, const fq& b) {
    uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 6; j++) { u128 s = (u128)a.l[j] * b.l[i] + t[j] + c; t[j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        u128 s = (u128)t[6] + c; t[6] = (uint64_t)s; t[7] = (uint64_t)(s >> 64);
        const uint64_t m = t[0] * INV;
        s = (u128)m * P[0] + t[0]; c = (uint64_t)(s >> 64);
        for (int j = 1; j < 6; j++) { s = (u128)m * P[j] + t[j] + c; t[j - 1] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        s = (u128)t[6] + c; t[5] = (uint64_t)s; t[6] = t[7] + (uint64_t)(s >> 64);
    }
    if (t[6] || geq_p(t)) {
        uint64_t br = 0;
        for (int i = 0; i < 6; i++) { u128 d = (u128)t[i] - P[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    }
    fq r; memcpy(r.l, t, 48); return r;
}
// a^-1 in Montgomery form for a in Montgomery form: plain inverse x = (a R)^-1 by binary GCD
// (host_inv.h, ~5 us instead of the ~37 us of a Fermat chain), then x R^3 R^-1 = R / a.
static inline fq inv(const fq& a) {
    static const fq R3 = [] {
        fq r2;
        for (int i = 0; i < 6; i++) r2.l[i] = (uint64_t)FqParams::r2(2 * i) | ((uint64_t)FqParams::r2(2 * i + 1) << 32);
        return mul(r2, r2);   // R^4 / R
    }();
    fq x;
    hostinv::inv_mod<6>(a.l, P, x.l);
    return mul(x, R3);
}


static inline fq add(const fq& a, const fq& b) {
    uint64_t t[6], c = 0;
    for (int i = 0; i < 6; i++) { u128 s = (u128)a.l[i] + b.l[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    if (c || geq_p(t)) {
        uint64_t br = 0;
        for (int i = 0; i < 6; i++) { u128 d = (u128)t[i] - P[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    }
    fq r; memcpy(r.l, t, 48); return r;
}
static inline fq sub(const fq& a, const fq& b) {
    uint64_t t[6], br = 0;
    for (int i = 0; i < 6; i++) { u128 d = (u128)a.l[i] - b.l[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    if (br) {
        uint64_t c = 0;
        for (int i = 0; i < 6; i++) { u128 s2 = (u128)t[i] + P[i] + c; t[i] = (uint64_t)s2; c = (uint64_t)(s2 >> 64); }
    }
    fq r; memcpy(r.l, t, 48); return r;
}
static inline bool is_zero(const fq& a) { uint64_t x = 0; for (int i = 0; i < 6; i++) x |= a.l[i]; return x == 0; }
}  // namespace hostfq


}  // namespace zkp

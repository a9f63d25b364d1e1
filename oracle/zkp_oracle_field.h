/*
 * Fr / Fq Montgomery arithmetic shared by the C restatement files -- TEST INFRASTRUCTURE.
 * 4 / 6 x u64 little-endian limbs, R = 2^256 / 2^384 (layout pinned by src/lib.rs:583-588).
 * PARITY UNPINNED: see zkp_oracle.c.
 */
#ifndef ZKP_ORACLE_FIELD_H
#define ZKP_ORACLE_FIELD_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ moduli */
static const uint64_t FR_P[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL,
                                 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t FR_INV = 0xfffffffeffffffffULL;
/* R mod r, R^2 mod r (R = 2^256) */
static const uint64_t FR_ONE[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL,
                                   0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL};
static const uint64_t FR_R2[4] = {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL,
                                  0x05d314967254398fULL, 0x0748d9d99f59ff11ULL};

static const uint64_t FQ_P[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL,
                                 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL,
                                 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t FQ_INV = 0x89f3fffcfffcfffdULL;
static const uint64_t FQ_ONE[6] = {0x760900000002fffdULL, 0xebf4000bc40c0002ULL,
                                   0x5f48985753c758baULL, 0x77ce585370525745ULL,
                                   0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};

/* ------------------------------------------------- generic N-limb helpers */
#define DEF_FIELD(PFX, N, MOD, INV)                                                     \
    typedef struct { uint64_t l[N]; } PFX##_t;                                          \
    static inline int PFX##_is_zero(const PFX##_t *a) {                                 \
        uint64_t x = 0; for (int i = 0; i < N; i++) x |= a->l[i]; return x == 0; }      \
    static inline int PFX##_eq(const PFX##_t *a, const PFX##_t *b) {                    \
        uint64_t x = 0; for (int i = 0; i < N; i++) x |= a->l[i] ^ b->l[i];             \
        return x == 0; }                                                                \
    static inline int PFX##_geq_p(const uint64_t *a) {                                  \
        for (int i = N - 1; i >= 0; i--) {                                              \
            if (a[i] > MOD[i]) return 1; if (a[i] < MOD[i]) return 0; }                 \
        return 1; }                                                                     \
    static inline void PFX##_sub_p(uint64_t *a) {                                       \
        uint64_t br = 0;                                                                \
        for (int i = 0; i < N; i++) {                                                   \
            u128 d = (u128)a[i] - MOD[i] - br; a[i] = (uint64_t)d;                      \
            br = (uint64_t)(d >> 64) & 1; } }                                           \
    static inline void PFX##_add(PFX##_t *r, const PFX##_t *a, const PFX##_t *b) {      \
        uint64_t c = 0;                                                                 \
        for (int i = 0; i < N; i++) {                                                   \
            u128 s = (u128)a->l[i] + b->l[i] + c; r->l[i] = (uint64_t)s;                \
            c = (uint64_t)(s >> 64); }                                                  \
        if (c || PFX##_geq_p(r->l)) PFX##_sub_p(r->l); }                                \
    static inline void PFX##_sub(PFX##_t *r, const PFX##_t *a, const PFX##_t *b) {      \
        uint64_t br = 0;                                                                \
        for (int i = 0; i < N; i++) {                                                   \
            u128 d = (u128)a->l[i] - b->l[i] - br; r->l[i] = (uint64_t)d;               \
            br = (uint64_t)(d >> 64) & 1; }                                             \
        if (br) { uint64_t c = 0;                                                       \
            for (int i = 0; i < N; i++) {                                               \
                u128 s = (u128)r->l[i] + MOD[i] + c; r->l[i] = (uint64_t)s;             \
                c = (uint64_t)(s >> 64); } } }                                          \
    static inline void PFX##_neg(PFX##_t *r, const PFX##_t *a) {                        \
        if (PFX##_is_zero(a)) { *r = *a; return; }                                      \
        uint64_t br = 0;                                                                \
        for (int i = 0; i < N; i++) {                                                   \
            u128 d = (u128)MOD[i] - a->l[i] - br; r->l[i] = (uint64_t)d;                \
            br = (uint64_t)(d >> 64) & 1; } }                                           \
    /* CIOS Montgomery multiplication */                                                \
    static inline void PFX##_mul(PFX##_t *r, const PFX##_t *a, const PFX##_t *b) {      \
        uint64_t t[N + 2]; memset(t, 0, sizeof t);                                      \
        for (int i = 0; i < N; i++) {                                                   \
            uint64_t c = 0;                                                             \
            for (int j = 0; j < N; j++) {                                               \
                u128 s = (u128)a->l[j] * b->l[i] + t[j] + c;                            \
                t[j] = (uint64_t)s; c = (uint64_t)(s >> 64); }                          \
            u128 s = (u128)t[N] + c; t[N] = (uint64_t)s; t[N + 1] = (uint64_t)(s >> 64);\
            uint64_t m = t[0] * INV;                                                    \
            s = (u128)m * MOD[0] + t[0]; c = (uint64_t)(s >> 64);                       \
            for (int j = 1; j < N; j++) {                                               \
                s = (u128)m * MOD[j] + t[j] + c;                                        \
                t[j - 1] = (uint64_t)s; c = (uint64_t)(s >> 64); }                      \
            s = (u128)t[N] + c; t[N - 1] = (uint64_t)s;                                 \
            t[N] = t[N + 1] + (uint64_t)(s >> 64); }                                    \
        if (t[N] || PFX##_geq_p(t)) PFX##_sub_p(t);                                     \
        memcpy(r->l, t, N * 8); }                                                       \
    static inline void PFX##_sqr(PFX##_t *r, const PFX##_t *a) { PFX##_mul(r, a, a); }

DEF_FIELD(fr, 4, FR_P, FR_INV)
DEF_FIELD(fq, 6, FQ_P, FQ_INV)

static inline void fr_from_mont(fr_t *r, const fr_t *a) {
    fr_t one = {{1, 0, 0, 0}};
    fr_mul(r, a, &one);
}
static inline void fr_to_mont(fr_t *r, const fr_t *a) {
    fr_t r2; memcpy(r2.l, FR_R2, 32);
    fr_mul(r, a, &r2);
}
static void fr_pow_u64(fr_t *r, const fr_t *a, uint64_t e) {
    fr_t acc; memcpy(acc.l, FR_ONE, 32);
    fr_t base = *a;
    while (e) {
        if (e & 1) fr_mul(&acc, &acc, &base);
        fr_sqr(&base, &base);
        e >>= 1;
    }
    *r = acc;
}
/* a^(r-2) */
static void fr_inv(fr_t *r, const fr_t *a) {
    uint64_t e[4]; memcpy(e, FR_P, 32); e[0] -= 2;
    fr_t acc; memcpy(acc.l, FR_ONE, 32);
    for (int i = 255; i >= 0; i--) {
        fr_sqr(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) fr_mul(&acc, &acc, a);
    }
    *r = acc;
}
static void fq_inv(fq_t *r, const fq_t *a) {
    uint64_t e[6]; memcpy(e, FQ_P, 48); e[0] -= 2;
    fq_t acc; memcpy(acc.l, FQ_ONE, 48);
    for (int i = 383; i >= 0; i--) {
        fq_sqr(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) fq_mul(&acc, &acc, a);
    }
    *r = acc;
}

static inline void fr_set_u64(fr_t *r, uint64_t v) {
    fr_t raw = {{v, 0, 0, 0}};
    fr_to_mont(r, &raw);
}
#endif

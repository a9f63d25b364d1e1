/*
 * CPU restatement of the zkplonk hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * PARITY UNPINNED: the reference (/root/reference) cannot be built here (no Rust) and
 * the crates holding its arithmetic (poly-commit, bls-12-381, zksnarks; path deps with
 * no pinned version, Cargo.toml:28-33) are absent; its tests hold no golden vector for
 * this path.  This file restates the published algorithms the reference reaches:
 *
 *   - Fr / Fq Montgomery arithmetic, 4 / 6 x u64 little-endian limbs, R = 2^256 / 2^384
 *     (layout pinned by src/lib.rs:583-588 MINUS_ONE);
 *   - poly_commit::Fft::{dft,idft,coset_dft,coset_idft}: bit-reverse + radix-2 DIT,
 *     coset shift g = 7, n^-1 scaling (call sites src/prover.rs:121-124,192,229;
 *     src/prover/quotient_poly.rs:54-58,115,145,237; src/key.rs:121-131,226-245);
 *   - poly_commit::msm_curve_addition / PlonkParams::commit: Pippenger bucket MSM
 *     (call sites src/prover.rs:133-136,194,262-265,440,452; src/key.rs:138-159);
 *   - the element-wise prover loops of src/prover/quotient_poly.rs:106-114,154-217,
 *     245-261 and src/permutation.rs:248-299 (see zkp_oracle_prover.c).
 *
 * It is validated against the dependency-free Python big-int oracle (oracle/*.py) by
 * tests/test_oracle_c.py and is the "port" CPU baseline timed by bench.py.
 * Threads: OpenMP over all host cores (the reference uses the rayon global pool).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#include "zkp_oracle_field.h"

/* exported single-op probes (used by tests to pin the C arithmetic to Python) */
void oracle_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    fr_mul((fr_t *)out, (const fr_t *)a, (const fr_t *)b);
}
void oracle_fq_mul(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    fq_mul((fq_t *)out, (const fq_t *)a, (const fq_t *)b);
}
void oracle_fr_inv(const uint64_t *a, uint64_t *out) { fr_inv((fr_t *)out, (const fr_t *)a); }
void oracle_fr_add(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    fr_add((fr_t *)out, (const fr_t *)a, (const fr_t *)b);
}
void oracle_fr_sub(const uint64_t *a, const uint64_t *b, uint64_t *out) {
    fr_sub((fr_t *)out, (const fr_t *)a, (const fr_t *)b);
}

/* ------------------------------------------------------------------- NTT */
/* 2^32-th root of unity 7^((r-1)/2^32), Montgomery form is computed at run time */
static const uint64_t FR_ROOT_RAW[4] = {0x3829971f439f0d2bULL, 0xb63683508c2280b9ULL,
                                        0xd09b681922c813b4ULL, 0x16a2a19edfe81f20ULL};

static void fr_domain_root(fr_t *w, unsigned k) {
    fr_t raw; memcpy(raw.l, FR_ROOT_RAW, 32);
    fr_to_mont(w, &raw);
    for (unsigned i = k; i < 32; i++) fr_sqr(w, w);
}

static inline size_t bitrev(size_t x, unsigned k) {
    size_t r = 0;
    for (unsigned i = 0; i < k; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* data: n = 2^k Montgomery Fr elements, in place.  len_in <= n real inputs, the rest
 * is treated as zero (Fft pads short inputs).  inverse: use w^-1 and scale by n^-1.
 * coset: forward -> multiply a_i by g^i first; inverse -> multiply by g^-i last. */
int oracle_ntt(uint64_t *data, size_t len_in, unsigned k, int inverse, int coset,
               int nthreads) {
    size_t n = (size_t)1 << k;
    if (len_in > n) return -1;
    fr_t *a = (fr_t *)data;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    for (size_t i = len_in; i < n; i++) memset(&a[i], 0, sizeof(fr_t));

    fr_t w, g, one; memcpy(one.l, FR_ONE, 32);
    fr_domain_root(&w, k);
    fr_set_u64(&g, 7);
    if (inverse) { fr_inv(&w, &w); fr_inv(&g, &g); }

    /* per-chunk geometric progressions for the coset shift */
    if (coset && !inverse) {
#pragma omp parallel num_threads(nthreads)
        {
            int t = omp_get_thread_num(), T = omp_get_num_threads();
            size_t lo = n * t / T, hi = n * (t + 1) / T;
            fr_t gi; fr_pow_u64(&gi, &g, lo);
            for (size_t i = lo; i < hi; i++) { fr_mul(&a[i], &a[i], &gi); fr_mul(&gi, &gi, &g); }
        }
    }
    /* bit-reverse */
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, k);
        if (i < j) { fr_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    /* half-size twiddle table */
    size_t half = n > 1 ? n / 2 : 1;
    fr_t *tw = (fr_t *)malloc(half * sizeof(fr_t));
    if (!tw) return -2;
#pragma omp parallel num_threads(nthreads)
    {
        int t = omp_get_thread_num(), T = omp_get_num_threads();
        size_t lo = half * t / T, hi = half * (t + 1) / T;
        if (lo < hi) {
            fr_t wi; fr_pow_u64(&wi, &w, lo);
            for (size_t i = lo; i < hi; i++) { tw[i] = wi; fr_mul(&wi, &wi, &w); }
        }
    }
    for (size_t m = 1; m < n; m <<= 1) {
        size_t stride = n / (2 * m);
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (size_t b = 0; b < n / 2; b++) {
            size_t j = b & (m - 1);
            size_t start = (b / m) * 2 * m;
            fr_t t, u = a[start + j];
            fr_mul(&t, &a[start + j + m], &tw[j * stride]);
            fr_add(&a[start + j], &u, &t);
            fr_sub(&a[start + j + m], &u, &t);
        }
    }
    free(tw);
    if (inverse) {
        fr_t ninv; fr_set_u64(&ninv, (uint64_t)n); fr_inv(&ninv, &ninv);
#pragma omp parallel num_threads(nthreads)
        {
            int t = omp_get_thread_num(), T = omp_get_num_threads();
            size_t lo = n * t / T, hi = n * (t + 1) / T;
            fr_t gi = ninv;
            if (coset) { fr_t gp; fr_pow_u64(&gp, &g, lo); fr_mul(&gi, &gi, &gp); }
            for (size_t i = lo; i < hi; i++) {
                fr_mul(&a[i], &a[i], &gi);
                if (coset) fr_mul(&gi, &gi, &g);
            }
        }
    }
    return 0;
}

/* -------------------------------------------------------------------- G1 */
typedef struct { fq_t x, y; } g1a_t;        /* affine, (0,0) = infinity (ABI convention) */
typedef struct { fq_t x, y, z; } g1j_t;     /* Jacobian, z = 0 infinity */

static inline int g1a_is_inf(const g1a_t *p) { return fq_is_zero(&p->x) && fq_is_zero(&p->y); }
static inline void g1j_set_inf(g1j_t *p) { memset(p, 0, sizeof *p); }

static void g1j_double(g1j_t *r, const g1j_t *p) {
    if (fq_is_zero(&p->z) || fq_is_zero(&p->y)) { g1j_set_inf(r); return; }
    fq_t A, B, C, D, E, F, t, Z3;
    fq_sqr(&A, &p->x); fq_sqr(&B, &p->y); fq_sqr(&C, &B);
    fq_add(&t, &p->x, &B); fq_sqr(&t, &t); fq_sub(&t, &t, &A); fq_sub(&t, &t, &C);
    fq_add(&D, &t, &t);
    fq_add(&E, &A, &A); fq_add(&E, &E, &A);
    fq_sqr(&F, &E);
    fq_mul(&Z3, &p->y, &p->z); fq_add(&Z3, &Z3, &Z3);
    fq_sub(&r->x, &F, &D); fq_sub(&r->x, &r->x, &D);
    fq_sub(&t, &D, &r->x); fq_mul(&t, &E, &t);
    fq_add(&C, &C, &C); fq_add(&C, &C, &C); fq_add(&C, &C, &C);
    fq_sub(&r->y, &t, &C);
    r->z = Z3;
}

static void g1j_add_affine(g1j_t *r, const g1j_t *p, const g1a_t *q) {
    if (g1a_is_inf(q)) { *r = *p; return; }
    if (fq_is_zero(&p->z)) { r->x = q->x; r->y = q->y; memcpy(r->z.l, FQ_ONE, 48); return; }
    fq_t Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
    fq_sqr(&Z1Z1, &p->z);
    fq_mul(&U2, &q->x, &Z1Z1);
    fq_mul(&S2, &q->y, &p->z); fq_mul(&S2, &S2, &Z1Z1);
    if (fq_eq(&U2, &p->x)) {
        if (fq_eq(&S2, &p->y)) { g1j_double(r, p); return; }
        g1j_set_inf(r); return;
    }
    fq_sub(&H, &U2, &p->x);
    fq_sqr(&HH, &H);
    fq_add(&I, &HH, &HH); fq_add(&I, &I, &I);
    fq_mul(&J, &H, &I);
    fq_sub(&rr, &S2, &p->y); fq_add(&rr, &rr, &rr);
    fq_mul(&V, &p->x, &I);
    fq_t X3, Y3, Z3;
    fq_sqr(&X3, &rr); fq_sub(&X3, &X3, &J); fq_sub(&X3, &X3, &V); fq_sub(&X3, &X3, &V);
    fq_sub(&t, &V, &X3); fq_mul(&Y3, &rr, &t);
    fq_mul(&t, &p->y, &J); fq_add(&t, &t, &t); fq_sub(&Y3, &Y3, &t);
    fq_add(&Z3, &p->z, &H); fq_sqr(&Z3, &Z3); fq_sub(&Z3, &Z3, &Z1Z1); fq_sub(&Z3, &Z3, &HH);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static void g1j_add(g1j_t *r, const g1j_t *p, const g1j_t *q) {
    if (fq_is_zero(&p->z)) { *r = *q; return; }
    if (fq_is_zero(&q->z)) { *r = *p; return; }
    fq_t Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
    fq_sqr(&Z1Z1, &p->z); fq_sqr(&Z2Z2, &q->z);
    fq_mul(&U1, &p->x, &Z2Z2); fq_mul(&U2, &q->x, &Z1Z1);
    fq_mul(&S1, &p->y, &q->z); fq_mul(&S1, &S1, &Z2Z2);
    fq_mul(&S2, &q->y, &p->z); fq_mul(&S2, &S2, &Z1Z1);
    if (fq_eq(&U1, &U2)) {
        if (fq_eq(&S1, &S2)) { g1j_double(r, p); return; }
        g1j_set_inf(r); return;
    }
    fq_sub(&H, &U2, &U1);
    fq_add(&I, &H, &H); fq_sqr(&I, &I);
    fq_mul(&J, &H, &I);
    fq_sub(&rr, &S2, &S1); fq_add(&rr, &rr, &rr);
    fq_mul(&V, &U1, &I);
    fq_t X3, Y3, Z3;
    fq_sqr(&X3, &rr); fq_sub(&X3, &X3, &J); fq_sub(&X3, &X3, &V); fq_sub(&X3, &X3, &V);
    fq_sub(&t, &V, &X3); fq_mul(&Y3, &rr, &t);
    fq_mul(&t, &S1, &J); fq_add(&t, &t, &t); fq_sub(&Y3, &Y3, &t);
    fq_add(&Z3, &p->z, &q->z); fq_sqr(&Z3, &Z3); fq_sub(&Z3, &Z3, &Z1Z1); fq_sub(&Z3, &Z3, &Z2Z2);
    fq_mul(&Z3, &Z3, &H);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static void g1j_to_affine(g1a_t *r, const g1j_t *p) {
    if (fq_is_zero(&p->z)) { memset(r, 0, sizeof *r); return; }
    fq_t zi, zi2;
    fq_inv(&zi, &p->z); fq_sqr(&zi2, &zi);
    fq_mul(&r->x, &p->x, &zi2);
    fq_mul(&zi2, &zi2, &zi);
    fq_mul(&r->y, &p->y, &zi2);
}

static inline unsigned get_window(const uint64_t *s, unsigned bit, unsigned c) {
    unsigned limb = bit / 64, off = bit % 64;
    uint64_t v = s[limb] >> off;
    if (off + c > 64 && limb + 1 < 4) v |= s[limb + 1] << (64 - off);
    return (unsigned)(v & (((uint64_t)1 << c) - 1));
}

/* Pippenger over one contiguous chunk, all windows; result Jacobian. */
static void msm_chunk(g1j_t *out, const g1a_t *bases, const fr_t *scalars_raw, size_t n,
                      unsigned c) {
    unsigned nwin = (255 + c - 1) / c;
    size_t nb = ((size_t)1 << c) - 1;
    g1j_t *buckets = (g1j_t *)malloc(nb * sizeof(g1j_t));
    g1j_t total; g1j_set_inf(&total);
    for (int w = (int)nwin - 1; w >= 0; w--) {
        for (unsigned d = 0; d < c; d++) g1j_double(&total, &total);
        memset(buckets, 0, nb * sizeof(g1j_t));
        for (size_t i = 0; i < n; i++) {
            unsigned d = get_window(scalars_raw[i].l, w * c, c);
            if (d) g1j_add_affine(&buckets[d - 1], &buckets[d - 1], &bases[i]);
        }
        g1j_t run, acc; g1j_set_inf(&run); g1j_set_inf(&acc);
        for (size_t b = nb; b-- > 0;) {
            g1j_add(&run, &run, &buckets[b]);
            g1j_add(&acc, &acc, &run);
        }
        g1j_add(&total, &total, &acc);
    }
    free(buckets);
    *out = total;
}

/* bases: n x 12 u64 (x,y Montgomery; (0,0) infinity); scalars: n x 4 u64 Montgomery Fr.
 * out: 12 u64 affine Montgomery ((0,0) if infinity).  Mirrors msm_curve_addition +
 * Commitment::new (projective -> affine). */
int oracle_msm_g1(const uint64_t *bases, const uint64_t *scalars, size_t n, uint64_t *out,
                  int nthreads) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    const g1a_t *B = (const g1a_t *)bases;
    fr_t *raw = (fr_t *)malloc((n ? n : 1) * sizeof(fr_t));
    if (!raw) return -2;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (size_t i = 0; i < n; i++) fr_from_mont(&raw[i], &((const fr_t *)scalars)[i]);

    size_t nchunks = (size_t)nthreads;
    if (nchunks > n / 32 + 1) nchunks = n / 32 + 1;
    size_t per = (n + nchunks - 1) / nchunks;
    unsigned c = 3;
    { size_t m = per; unsigned lg = 0; while (m > 1) { m >>= 1; lg++; }
      if (lg > 6) c = lg - 3; if (c > 16) c = 16; }
    g1j_t *parts = (g1j_t *)malloc(nchunks * sizeof(g1j_t));
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (size_t t = 0; t < nchunks; t++) {
        size_t lo = t * per, hi = lo + per; if (hi > n) hi = n; if (lo > hi) lo = hi;
        msm_chunk(&parts[t], B + lo, raw + lo, hi - lo, c);
    }
    g1j_t total; g1j_set_inf(&total);
    for (size_t t = 0; t < nchunks; t++) g1j_add(&total, &total, &parts[t]);
    g1j_to_affine((g1a_t *)out, &total);
    free(parts); free(raw);
    return 0;
}

/* [s_i * G]: fixed-base multiples of an affine point; used to synthesise SRS-shaped
 * bases P_i = tau^i * G for full-size tests.  scalars raw (non-Montgomery). */
int oracle_g1_fixed_base_mul(const uint64_t *base, const uint64_t *scalars_raw, size_t n,
                             uint64_t *out, int nthreads) {
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    /* 8-bit window table: tbl[w][d-1] = d * 2^(8w) * base */
    enum { W = 32, D = 255 };
    g1a_t *tbl = (g1a_t *)malloc(sizeof(g1a_t) * W * D);
    g1j_t cur; cur.x = ((const g1a_t *)base)->x; cur.y = ((const g1a_t *)base)->y;
    memcpy(cur.z.l, FQ_ONE, 48);
    for (int w = 0; w < W; w++) {
        g1j_t acc; g1j_set_inf(&acc);
        for (int d = 0; d < D; d++) {
            g1j_add(&acc, &acc, &cur);
            g1j_to_affine(&tbl[w * D + d], &acc);
        }
        for (int i = 0; i < 8; i++) g1j_double(&cur, &cur);
    }
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (size_t i = 0; i < n; i++) {
        const uint64_t *s = scalars_raw + 4 * i;
        g1j_t acc; g1j_set_inf(&acc);
        for (int w = 0; w < W; w++) {
            unsigned d = (unsigned)((s[w / 8] >> (8 * (w % 8))) & 255);
            if (d) g1j_add_affine(&acc, &acc, &tbl[w * D + d - 1]);
        }
        g1j_to_affine((g1a_t *)(out + 12 * i), &acc);
    }
    free(tbl);
    return 0;
}

int oracle_num_threads(void) { return omp_get_max_threads(); }

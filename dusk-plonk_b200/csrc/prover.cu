// Element-wise prover rounds on device-resident polynomials (sm_100a).
//
// Replaces the host loops the reference runs between its NTTs and commits:
//   * Permutation::compute_permutation_vec           src/permutation.rs:205-300
//   * quotient_poly::compute (gate + permutation identities, / Z_H)
//                                                    src/prover/quotient_poly.rs:20-118,122-262
//   * Coefficients::evaluate (16 openings + t, r)    src/prover/linearization_poly.rs:52-73,108
//   * widget.linearize sums and `&poly * &scalar + ..` src/prover/linearization_poly.rs:75-105,
//                                                    src/prover.rs:408-418
//   * PlonkParams::compute_aggregate_witness          src/prover.rs:422-451
//   * Coefficients::blind                             src/prover.rs:126-129,193
//   * compute_permutation_lagrange                    src/permutation.rs:140-169
// Every result is an exact function of its inputs over Fr, so any evaluation order gives
// the reference's bits; the kernels are organised for the GPU, not after the Rust loops:
//   - n per-gate inversions + a serial product scan  ->  two chunked product scans + ONE inversion
//   - 8n inversions of Z_H (8 distinct values)        ->  an 8-entry table
//   - 17 serial Horner passes                         ->  batched chunk-Horner + power-table tree
//   - synthetic division                              ->  chunked suffix-Horner scan
#include <string.h>

#include "common.cuh"
#include "host_inv.h"

namespace zkp {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ fr_t pld(const fr_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void pst(fr_t* p, const fr_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// x * K for a small compile-time K by double-and-add (additions only)
template <unsigned K>
__device__ __forceinline__ fr_t mul_small(const fr_t& x) {
    static_assert(K >= 1, "K >= 1");
    if constexpr (K == 1) {
        return x;
    } else {
        fr_t h = dbl(mul_small<(K >> 1)>(x));
        if constexpr (K & 1) h = h + x;
        return h;
    }
}

static inline fr_t fr_from_host(const uint64_t* p) { fr_t r; memcpy(r.l, p, 32); return r; }

#define CHECK_REF(r) ((r).buf && (r).off + (r).len <= (r).buf->n)

// ------------------------------------------------------------------ small kernels
__global__ void fill_kernel(fr_t* out, size_t n, fr_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pst(out + i, v);
}

// Coefficients::blind: p <- p + (b0 + b1 X + ..)(X^n - 1)
__global__ void blind_kernel(fr_t* p, size_t n, fr_t b0, fr_t b1, fr_t b2, unsigned cnt) {
    unsigned i = threadIdx.x;
    if (i >= cnt) return;
    fr_t b = i == 0 ? b0 : (i == 1 ? b1 : b2);
    pst(p + i, pld(p + i) - b);
    pst(p + n + i, b);
}

// sigma evaluations: out[i] = K[wire] * w^gate, enc = wire << 30 | gate
__global__ void perm_lagrange_kernel(const uint32_t* enc, size_t n, const fr_t* roots, fr_t k1, fr_t k2, fr_t k3,
                                     fr_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = enc[i];
    const uint32_t w = e >> 30;
    fr_t r = pld(roots + (e & 0x3fffffffu));
    if (w == 1) r = r * k1;
    else if (w == 2) r = r * k2;
    else if (w == 3) r = r * k3;
    pst(out + i, r);
}

// ------------------------------------------------------------------ permutation accumulator
struct PermArgs {
    const fr_t* w[4];
    const fr_t* sigma[4];
    const fr_t* roots;
    fr_t beta, gamma, bk[4];
    size_t n;
    fr_t* num;
    fr_t* den;
};

__global__ void __launch_bounds__(128) perm_numden_kernel(PermArgs a) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const fr_t br = a.beta * pld(a.roots + i);  // beta w^i; beta K_j w^i by additions (K = 1, 7, 13, 17)
    const fr_t bkr[4] = {br, mul_small<7>(br), mul_small<13>(br), mul_small<17>(br)};
    fr_t nu, de;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const fr_t w = pld(a.w[j] + i) + a.gamma;
        const fr_t f = w + bkr[j];
        const fr_t g = w + a.beta * pld(a.sigma[j] + i);
        nu = j ? nu * f : f;
        de = j ? de * g : g;
    }
    pst(a.num + i, nu);
    pst(a.den + i, de);
}

// Chunked product scan, three passes.  reverse = 0: inclusive prefix products;
// reverse = 1: inclusive suffix products.
static constexpr unsigned SCAN_L = 16;

__global__ void prod_chunk_kernel(const fr_t* in, size_t n, fr_t* chunk) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * SCAN_L;
    if (lo >= n) return;
    size_t hi = lo + SCAN_L < n ? lo + SCAN_L : n;
    fr_t p = pld(in + lo);
    for (size_t i = lo + 1; i < hi; i++) p = p * pld(in + i);
    pst(chunk + t, p);
}

// one block per scan: block 0 turns fwd[t] into the product of the chunks before t, block 1 turns
// rev[t] into the product of the chunks after t (the numerator and denominator scans of the
// permutation accumulator are independent: one launch, two SMs)
__global__ void __launch_bounds__(1024) prod_carry_kernel(fr_t* fwd, fr_t* rev, size_t nc) {
    __shared__ fr_t sm[1024];
    const unsigned T = blockDim.x, tid = threadIdx.x;
    const int reverse = blockIdx.x;
    fr_t* chunk = reverse ? rev : fwd;
    const size_t per = (nc + T - 1) / T;
    const unsigned active = (unsigned)((nc + per - 1) / per);   // threads that own chunks
    // logical index j runs in scan direction
    auto at = [&](size_t j) { return reverse ? nc - 1 - j : j; };
    const size_t lo = (size_t)tid * per, hi = lo + per < nc ? lo + per : nc;
    fr_t p = fr_t::one();
    for (size_t j = lo; j < hi; j++) p = p * pld(chunk + at(j));
    sm[tid] = p;
    __syncthreads();
    for (unsigned s = 1; s < active; s <<= 1) {   // threads beyond `active` hold one: no level needed for them
        fr_t v = sm[tid];
        if (tid >= s) v = sm[tid - s] * v;
        __syncthreads();
        sm[tid] = v;
        __syncthreads();
    }
    fr_t run = tid ? sm[tid - 1] : fr_t::one();
    for (size_t j = lo; j < hi; j++) {
        fr_t c = pld(chunk + at(j));
        pst(chunk + at(j), run);
        run = run * c;
    }
}

__global__ void prod_apply_kernel(const fr_t* in, size_t n, const fr_t* carry, int reverse, fr_t* out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * SCAN_L;
    if (lo >= n) return;
    size_t hi = lo + SCAN_L < n ? lo + SCAN_L : n;
    fr_t run = pld(carry + t);
    if (!reverse) {
        for (size_t i = lo; i < hi; i++) { run = run * pld(in + i); pst(out + i, run); }
    } else {
        for (size_t i = hi; i-- > lo;) { run = run * pld(in + i); pst(out + i, run); }
    }
}

// Host-side Fr inversion (4 x u64 CIOS Montgomery, Fermat): the one inversion of the permutation
// accumulator would keep a single GPU thread busy for ~0.2 ms.
namespace hostfr {
typedef unsigned __int128 u128;
static const uint64_t P[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                              0x73eda753299d7d48ULL};
static const uint64_t INV = 0xfffffffeffffffffULL;
static const uint64_t ONE[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL,
                                0x1824b159acc5056fULL};
struct fr { uint64_t l[4]; };
static inline fr mul(const fr& a, const fr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) { u128 s = (u128)a.l[j] * b.l[i] + t[j] + c; t[j] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        u128 s = (u128)t[4] + c; t[4] = (uint64_t)s; t[5] = (uint64_t)(s >> 64);
        const uint64_t m = t[0] * INV;
        s = (u128)m * P[0] + t[0]; c = (uint64_t)(s >> 64);
        for (int j = 1; j < 4; j++) { s = (u128)m * P[j] + t[j] + c; t[j - 1] = (uint64_t)s; c = (uint64_t)(s >> 64); }
        s = (u128)t[4] + c; t[3] = (uint64_t)s; t[4] = t[5] + (uint64_t)(s >> 64);
    }
    bool ge = t[4] != 0;
    if (!ge) {
        ge = true;
        for (int i = 3; i >= 0; i--) { if (t[i] > P[i]) break; if (t[i] < P[i]) { ge = false; break; } }
    }
    if (ge) {
        uint64_t br = 0;
        for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - P[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    }
    fr r; memcpy(r.l, t, 32); return r;
}
// a^-1 in Montgomery form: plain inverse of (a R) by binary GCD (host_inv.h), times R^3 R^-1
static inline fr inv(const fr& a) {
    static const fr R3 = [] {
        fr r2;
        for (int i = 0; i < 4; i++) r2.l[i] = (uint64_t)FrParams::r2(2 * i) | ((uint64_t)FrParams::r2(2 * i + 1) << 32);
        return mul(r2, r2);
    }();
    fr x;
    hostinv::inv_mod<4>(a.l, P, x.l);
    return mul(x, R3);
}
}  // namespace hostfr

// z[0] = 1, z[i] = P[i-1] * S[i] * inv   (P inclusive prefix of num, S inclusive suffix of den,
// inv = 1 / prod(den)):  prod_{j<i} num_j / den_j
__global__ void perm_z_combine_kernel(const fr_t* P, const fr_t* S, fr_t inv, size_t n, fr_t* z) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) { pst(z, fr_t::one()); return; }
    pst(z + i, pld(P + i - 1) * pld(S + i) * inv);
}

// ------------------------------------------------------------------ quotient
struct QuotArgs {
    const fr_t* w[4];
    const fr_t* z;
    const fr_t* pi;
    const fr_t* l1;
    const fr_t* sel[11];  // q_m q_l q_r q_o q_c q_4 q_arith q_range q_logic q_fixed q_var
    const fr_t* sigma[4];
    const fr_t* linear;
    fr_t alpha, beta, gamma, rs, ls, fs, vs;
    fr_t bk1, bk2, bk3;  // beta * K_j
    fr_t edwards_d;
    fr_t zh_inv[8];
    uint32_t mask;
    size_t n8;
    size_t first, count;  // evaluate indices [first, first + count) (a rank's slice when sharded)
    int sliced;           // w / z / pi / l1 hold only [first, first + count + 8): index them relative to first
    unsigned coset_k;     // non-zero: coset layout (whole cosets of n = 2^coset_k points, see zkp_quotient_args)
    unsigned coset_first;
    fr_t* out;
};

__device__ __forceinline__ fr_t delta4(const fr_t& f, const fr_t& one, const fr_t& two, const fr_t& three) {
    return f * (f - one) * ((f - two) * (f - three));
}

__global__ void __launch_bounds__(128) quotient_kernel(const __grid_constant__ QuotArgs q) {
    const size_t t_ = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t_ >= q.count) return;
    const size_t i = q.first + t_;
    // "next gate" on the 8n coset (quotient_poly.rs:60-66).  wi / win index the witness-side vectors
    // (wires, z, PI, L1): absolute, or relative to the slice (whose 8-element halo holds the wrap-around);
    // in the coset layout the next gate is the next point of the same coset
    const size_t cmask = ((size_t)1 << q.coset_k) - 1;
    const size_t wi = q.sliced ? t_ : i;
    const size_t win = q.coset_k ? ((i & ~cmask) | ((i + 1) & cmask)) : q.sliced ? t_ + 8 : (i + 8) & (q.n8 - 1);
    const fr_t one = fr_t::one(), two = dbl(one), three = two + one;
    const fr_t a = pld(q.w[0] + wi), b = pld(q.w[1] + wi), c = pld(q.w[2] + wi), d = pld(q.w[3] + wi);
    const fr_t qc = pld(q.sel[4] + i);
    // arithmetic widget + PI
    fr_t t = a * b * pld(q.sel[0] + i) + a * pld(q.sel[1] + i) + b * pld(q.sel[2] + i) + c * pld(q.sel[3] + i) +
             d * pld(q.sel[5] + i) + qc;
    t = t * pld(q.sel[6] + i) + pld(q.pi + wi);
    if (q.mask) {
        const fr_t an = pld(q.w[0] + win), bn = pld(q.w[1] + win), dn = pld(q.w[3] + win);
        if (q.mask & 1u) {  // range
            const fr_t k = sqr(q.rs), k2 = sqr(k), k3 = k2 * k;
            fr_t s = delta4(c - mul_small<4>(d), one, two, three) + delta4(b - mul_small<4>(c), one, two, three) * k +
                     delta4(a - mul_small<4>(b), one, two, three) * k2 +
                     delta4(dn - mul_small<4>(a), one, two, three) * k3;
            t = t + s * q.rs * pld(q.sel[7] + i);
        }
        if (q.mask & 2u) {  // logic
            const fr_t k = sqr(q.ls), k2 = sqr(k), k3 = k2 * k, k4 = k3 * k;
            const fr_t A = an - mul_small<4>(a), B = bn - mul_small<4>(b), D = dn - mul_small<4>(d);
            const fr_t ab = A + B;
            // F = w (w (4w - 18(A+B) + 81) + 18(A^2 + B^2) - 81(A+B) + 83)
            fr_t f = mul_small<4>(c) - mul_small<18>(ab) + mul_small<81>(one);
            f = c * f + mul_small<18>(sqr(A) + sqr(B)) - mul_small<81>(ab) + mul_small<83>(one);
            f = c * f;
            const fr_t e = mul_small<3>(ab + D) - dbl(f);
            const fr_t bb = qc * (mul_small<9>(D) - mul_small<3>(ab));
            fr_t s = (c - A * B) * k3 + delta4(A, one, two, three) + delta4(B, one, two, three) * k +
                     delta4(D, one, two, three) * k2 + (bb + e) * k4;
            t = t + s * q.ls * pld(q.sel[8] + i);
        }
        if (q.mask & 4u) {  // fixed-base scalar mul
            const fr_t k = sqr(q.fs), k2 = sqr(k), k3 = k2 * k;
            const fr_t xb = pld(q.sel[1] + i), yb = pld(q.sel[2] + i);
            const fr_t bit = dn - dbl(d);
            const fr_t bitc = bit * (bit - one) * (bit + one);
            const fr_t ya = sqr(bit) * (yb - one) + one;
            const fr_t xa = bit * xb;
            const fr_t xyc = (bit * qc - c) * k;
            const fr_t tt = c * a * b * q.edwards_d;
            const fr_t xacc = ((an + an * tt) - (a * ya + b * xa)) * k2;
            const fr_t yacc = ((bn - bn * tt) - (b * ya + a * xa)) * k3;
            t = t + (bitc + xacc + yacc + xyc) * q.fs * pld(q.sel[9] + i);
        }
        if (q.mask & 8u) {  // variable-base addition
            const fr_t k = sqr(q.vs);
            const fr_t y1x2 = b * c, y1y2 = b * d, x1x2 = a * c;
            const fr_t tt = q.edwards_d * dn * y1x2;
            const fr_t xyc = a * d - dn;
            const fr_t x3c = ((dn + y1x2) - (an + an * tt)) * k;
            const fr_t y3c = ((y1y2 + x1x2) - (bn - bn * tt)) * sqr(k);
            t = t + (xyc + x3c + y3c) * q.vs * pld(q.sel[10] + i);
        }
    }
    // permutation argument (quotient_poly.rs:245-261)
    {
        const fr_t z = pld(q.z + wi), zn = pld(q.z + win), x = pld(q.linear + i);
        const fr_t ag = a + q.gamma, bg = b + q.gamma, cg = c + q.gamma, dg = d + q.gamma;
        // beta K_j x = K_j (beta x) with K = 7, 13, 17 (src/permutation.rs:28-30): additions, not multiplications
        const fr_t bx = q.beta * x;
        fr_t ident = (ag + bx) * (bg + mul_small<7>(bx)) * ((cg + mul_small<13>(bx)) * (dg + mul_small<17>(bx))) * z;
        fr_t copy = (ag + q.beta * pld(q.sigma[0] + i)) * (bg + q.beta * pld(q.sigma[1] + i)) *
                    ((cg + q.beta * pld(q.sigma[2] + i)) * (dg + q.beta * pld(q.sigma[3] + i))) * zn;
        t = t + (ident - copy) * q.alpha + (z - one) * pld(q.l1 + wi);
    }
    pst(q.out + i, t * q.zh_inv[q.coset_k ? ((q.coset_first + (i >> q.coset_k)) & 7) : (i & 7)]);
}

// ------------------------------------------------------------------ batched evaluation
static constexpr unsigned EV_MAX = 32;   // polynomials per launch
static constexpr unsigned EV_L = 32;     // coefficients per thread (the serial Horner part runs with every lane
                                         // busy; the tree that follows does not, so it is kept a small share)
static constexpr unsigned EV_T = 256;    // threads per block
static constexpr unsigned EV_BLOCK_LOG = 13;  // log2(EV_T * EV_L): a block's chunk is weighed by X^(2^13 * block)
static_assert((1u << EV_BLOCK_LOG) == EV_T * EV_L, "block chunk");

struct EvalArgs {
    const fr_t* p[EV_MAX];
    uint64_t len[EV_MAX];
    fr_t pw[2][28];        // point_k^(2^j) for the (up to) two evaluation points of a launch
    uint8_t which[EV_MAX]; // point of polynomial y
    unsigned nblocks;
    fr_t* partial;  // [count][nblocks]
};

// block b of polynomial y: value of coefficients [b*2048, (b+1)*2048) relative to the block start
__global__ void __launch_bounds__(EV_T) eval_block_kernel(const __grid_constant__ EvalArgs a) {
    __shared__ fr_t sm[EV_T];
    const unsigned y = blockIdx.y, tid = threadIdx.x;
    const fr_t* p = a.p[y];
    const size_t len = a.len[y];
    const fr_t* pw = a.pw[a.which[y]];
    // thread t takes coefficients base + t + j EV_T (consecutive threads read consecutive elements:
    // coalesced), Horner in X^EV_T; the tree then weighs thread t by X^t
    static_assert(EV_T == 256, "X^EV_T is pw[8]");
    const size_t base = (size_t)blockIdx.x * EV_T * EV_L;
    if (base >= len) {   // block beyond this (shorter) polynomial: nothing to add, skip the tree
        if (tid == 0) pst(a.partial + (size_t)y * a.nblocks + blockIdx.x, fr_t::zero());
        return;
    }
    fr_t s = fr_t::zero();
    {
#pragma unroll 4
        for (int j = EV_L - 1; j >= 0; j--) {
            const size_t i = base + (size_t)j * EV_T + tid;
            const fr_t c = i < len ? pld(p + i) : fr_t::zero();
            s = j == (int)EV_L - 1 ? c : s * pw[8] + c;
        }
    }
    sm[tid] = s;
    __syncthreads();
    unsigned lvl = 0;
    for (unsigned st = 1; st < EV_T; st <<= 1, lvl++) {
        if ((tid & (2 * st - 1)) == 0) sm[tid] = sm[tid] + pw[lvl] * sm[tid + st];
        __syncthreads();
    }
    if (tid == 0) pst(a.partial + (size_t)y * a.nblocks + blockIdx.x, sm[0]);
}

// one block per polynomial: sum_b partial[b] * X^b with X = point^2048
__global__ void __launch_bounds__(EV_T) eval_final_kernel(const __grid_constant__ EvalArgs a, fr_t* out) {
    __shared__ fr_t sm[EV_T];
    const unsigned y = blockIdx.x, tid = threadIdx.x;
    const fr_t* part = a.partial + (size_t)y * a.nblocks;
    const fr_t* pw = a.pw[a.which[y]];
    // thread t owns blocks [t*per, (t+1)*per): per is a power of two so X^per is in the table
    unsigned per_log = 0;
    while (((size_t)EV_T << per_log) < a.nblocks) per_log++;
    const size_t per = (size_t)1 << per_log;
    const size_t lo = (size_t)tid * per;
    fr_t s = fr_t::zero();
    if (lo < a.nblocks) {
        const size_t hi = lo + per < a.nblocks ? lo + per : a.nblocks;
        s = pld(part + hi - 1);
        for (size_t i = hi - 1; i-- > lo;) s = s * pw[EV_BLOCK_LOG] + pld(part + i);
    }
    sm[tid] = s;
    __syncthreads();
    unsigned lvl = EV_BLOCK_LOG + per_log;
    for (unsigned st = 1; st < EV_T; st <<= 1, lvl++) {
        if ((tid & (2 * st - 1)) == 0) sm[tid] = sm[tid] + pw[lvl < 28 ? lvl : 27] * sm[tid + st];
        __syncthreads();
    }
    if (tid == 0) pst(out + y, sm[0]);
}

// ------------------------------------------------------------------ linear combination
static constexpr unsigned LC_MAX = 16;
struct LincombArgs {
    const fr_t* p[LC_MAX];
    uint64_t len[LC_MAX];
    fr_t s[LC_MAX];
    unsigned count;
    size_t n;
    fr_t* out;
};

__global__ void __launch_bounds__(256) lincomb_kernel(const __grid_constant__ LincombArgs a) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    fr_t acc = fr_t::zero();
    for (unsigned k = 0; k < a.count; k++)
        if (i < a.len[k]) acc = acc + a.s[k] * pld(a.p[k] + i);
    pst(a.out + i, acc);
}

// ------------------------------------------------------------------ division by (X - point)
// h_j = sum_{i >= j} c_i point^(i-j);  quotient q_j = h_{j+1}  (ruffini, remainder dropped)
static constexpr unsigned DV_L = 16;

__global__ void div_chunk_kernel(const fr_t* c, size_t n, fr_t point, fr_t* chunk) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * DV_L;
    if (lo >= n) return;
    size_t hi = lo + DV_L < n ? lo + DV_L : n;
    fr_t s = pld(c + hi - 1);
    for (size_t i = hi - 1; i-- > lo;) s = s * point + pld(c + i);
    pst(chunk + t, s);
}

// Block b owns the chunks [b * tile, (b + 1) * tile) (the whole array when tile >= nc):
// chunk[t] <- sum_{t' > t, t' in the tile} chunk[t'] X^(t'-t-1)   (carry into chunk t from inside its tile)
// tile_val[b] <- value of the whole tile relative to its lowest chunk (if tile_val != nullptr).
// Two levels (tiles, then the same kernel over tile_val with X^tile) keep every block short.
static constexpr size_t DV_TILE = 2048;

__global__ void __launch_bounds__(512) div_carry_kernel(fr_t* chunk, size_t nc_all, size_t tile, fr_t X,
                                                       fr_t* tile_val) {
    __shared__ fr_t sm[512];
    __shared__ fr_t xp[512];
    const unsigned T = blockDim.x, tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * tile;
    if (base >= nc_all) return;
    const size_t nc = nc_all - base < tile ? nc_all - base : tile;
    chunk += base;
    const size_t per = (nc + T - 1) / T;
    const unsigned active = (unsigned)((nc + per - 1) / per);   // threads that own chunks
    // reversed logical order: j = 0 is the top chunk
    const size_t lo = (size_t)tid * per, hi = lo + per < nc ? lo + per : nc;
    // local value of this thread's span relative to the span's lowest chunk, and X^(span length)
    fr_t s = fr_t::zero(), xl = fr_t::one();
    for (size_t j = lo; j < hi; j++) { s = s * X + pld(chunk + (nc - 1 - j)); xl = xl * X; }
    sm[tid] = s;
    xp[tid] = xl;
    __syncthreads();
    // inclusive scan over threads of the affine maps carry -> carry * xl + s
    for (unsigned st = 1; st < active; st <<= 1) {   // idle threads hold the identity map
        fr_t v = sm[tid], w = xp[tid];
        if (tid >= st) { v = sm[tid - st] * w + v; w = xp[tid - st] * w; }
        __syncthreads();
        sm[tid] = v;
        xp[tid] = w;
        __syncthreads();
    }
    fr_t run = tid ? sm[tid - 1] : fr_t::zero();
    for (size_t j = lo; j < hi; j++) {
        const size_t idx = nc - 1 - j;
        fr_t cv = pld(chunk + idx);
        pst(chunk + idx, run);
        run = run * X + cv;
    }
    if (tile_val && lo < hi && hi == nc) pst(tile_val + blockIdx.x, run);
}

// carry: in-tile carries; tile_carry (may be null): carry into each tile from the tiles above.
__global__ void div_apply_kernel(const fr_t* c, size_t n, fr_t point, const fr_t* carry, const fr_t* tile_carry,
                                 size_t nc, fr_t X, fr_t* out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * DV_L;
    if (lo >= n) return;
    size_t hi = lo + DV_L < n ? lo + DV_L : n;
    fr_t h = pld(carry + t);  // h_{hi} from inside the tile
    if (tile_carry) {
        const size_t b = t / DV_TILE;
        const size_t tile_end = (b + 1) * DV_TILE < nc ? (b + 1) * DV_TILE : nc;
        h = h + pld(tile_carry + b) * pow_u64(X, (uint64_t)(tile_end - 1 - t));
    }
    for (size_t i = hi; i-- > lo;) {
        if (i + 1 < n) pst(out + i, h);  // q_i = h_{i+1}
        h = h * point + pld(c + i);
    }
}

// ------------------------------------------------------------------ scratch
static int scratch(zkp_ctx* ctx, size_t n, fr_t** out) {
    if (ctx->prover_scratch_n < n) {
        if (ctx->prover_scratch) cudaFree(ctx->prover_scratch);
        ctx->prover_scratch = nullptr;
        ctx->prover_scratch_n = 0;
        ZKP_CUDA(ctx, cudaMalloc(&ctx->prover_scratch, n * sizeof(fr_t)));
        ctx->prover_scratch_n = n;
    }
    *out = ctx->prover_scratch;
    return ZKP_OK;
}

void prover_free(zkp_ctx* ctx) {
    if (ctx->prover_scratch) cudaFree(ctx->prover_scratch);
    ctx->prover_scratch = nullptr;
    ctx->prover_scratch_n = 0;
}

static inline unsigned blocks_for(size_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

}  // namespace zkp

using namespace zkp;

extern "C" {

int zkp_ntt_ref_dev(zkp_ctx* ctx, zkp_poly_ref in, zkp_buf* out, size_t out_off, unsigned k, int inverse,
                    int coset) {
    if (!ctx || !out || !CHECK_REF(in) || k > 28) return ZKP_ERR_INVALID;
    const size_t n = (size_t)1 << k;
    if (in.len > n || out_off + n > out->n) return ZKP_ERR_INVALID;
    return ntt_run(ctx, in.buf->d + in.off, 0, in.len, out->d + out_off, 0, k, inverse != 0, coset != 0, 1);
}

int zkp_buf_fill(zkp_ctx* ctx, zkp_buf* buf, size_t off, size_t n, const uint64_t value[4]) {
    if (!ctx || !buf || !value || off + n > buf->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (!n) return ZKP_OK;
    fill_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(buf->d + off, n, fr_from_host(value));
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

int zkp_poly_blind_dev(zkp_ctx* ctx, zkp_buf* buf, size_t off, size_t n, const uint64_t* blinders, unsigned count) {
    if (!ctx || !buf || !blinders || count == 0 || count > 3 || off + n + count > buf->n || count > n)
        return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    fr_t b[3] = {fr_t::zero(), fr_t::zero(), fr_t::zero()};
    for (unsigned i = 0; i < count; i++) b[i] = fr_from_host(blinders + 4 * i);
    blind_kernel<<<1, 32, 0, ctx->stream>>>(buf->d + off, n, b[0], b[1], b[2], count);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

int zkp_perm_lagrange_dev(zkp_ctx* ctx, unsigned k, const uint32_t* enc, size_t n, const zkp_buf* roots,
                          zkp_buf* out, size_t out_off) {
    if (!ctx || !enc || !roots || !out || k > 28 || n > ((size_t)1 << k) || roots->n < ((size_t)1 << k) ||
        out_off + n > out->n || k > 30)
        return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    uint32_t* d_enc = nullptr;
    ZKP_CUDA(ctx, cudaMalloc(&d_enc, n * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpyAsync(d_enc, enc, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { cudaFree(d_enc); return cuda_fail(ctx, e, "h2d", __FILE__, __LINE__); }
    perm_lagrange_kernel<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(d_enc, n, roots->d, from_u64<FrParams>(7),
                                                                     from_u64<FrParams>(13), from_u64<FrParams>(17),
                                                                     out->d + out_off);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_enc);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "perm_lagrange", __FILE__, __LINE__);
    return ZKP_OK;
}

int zkp_perm_z_dev(zkp_ctx* ctx, size_t n, const zkp_poly_ref wires[4], const zkp_poly_ref sigmas[4],
                   const zkp_buf* roots, const uint64_t beta[4], const uint64_t gamma[4], zkp_buf* out,
                   size_t out_off) {
    if (!ctx || !wires || !sigmas || !roots || !beta || !gamma || !out || n == 0 || roots->n < n ||
        out_off + n > out->n)
        return ZKP_ERR_INVALID;
    for (int j = 0; j < 4; j++)
        if (!CHECK_REF(wires[j]) || !CHECK_REF(sigmas[j]) || wires[j].len < n || sigmas[j].len < n)
            return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    const size_t nc = (n + SCAN_L - 1) / SCAN_L;
    fr_t* s;
    if ((rc = scratch(ctx, 4 * n + 2 * nc + 1, &s))) return rc;
    fr_t *num = s, *den = s + n, *P = s + 2 * n, *S = s + 3 * n, *c1 = s + 4 * n, *c2 = c1 + nc, *inv = c2 + nc;
    PermArgs a;
    for (int j = 0; j < 4; j++) {
        a.w[j] = wires[j].buf->d + wires[j].off;
        a.sigma[j] = sigmas[j].buf->d + sigmas[j].off;
    }
    a.roots = roots->d;
    a.beta = fr_from_host(beta);
    a.gamma = fr_from_host(gamma);
    a.bk[0] = a.beta;
    a.bk[1] = a.beta * from_u64<FrParams>(7);
    a.bk[2] = a.beta * from_u64<FrParams>(13);
    a.bk[3] = a.beta * from_u64<FrParams>(17);
    a.n = n; a.num = num; a.den = den;
    cudaStream_t st = ctx->stream;
    ProfScope prof(ctx, "perm_z");
    perm_numden_kernel<<<blocks_for(n, 128), 128, 0, st>>>(a);
    ZKP_LAUNCHED(ctx);
    const unsigned cb = blocks_for(nc, 128);
    prod_chunk_kernel<<<cb, 128, 0, st>>>(num, n, c1);
    ZKP_LAUNCHED(ctx);
    prod_chunk_kernel<<<cb, 128, 0, st>>>(den, n, c2);
    ZKP_LAUNCHED(ctx);
    // few warps per scheduler keep each thread's multiplication chain at its latency, not the pipe's
    prod_carry_kernel<<<2, nc <= 16384 ? 256 : 1024, 0, st>>>(c1, c2, nc);
    ZKP_LAUNCHED(ctx);
    prod_apply_kernel<<<cb, 128, 0, st>>>(num, n, c1, 0, P);
    ZKP_LAUNCHED(ctx);
    prod_apply_kernel<<<cb, 128, 0, st>>>(den, n, c2, 1, S);
    ZKP_LAUNCHED(ctx);
    // 1 / prod(den) on the host: S[0] is the product of every denominator
    hostfr::fr* hp = reinterpret_cast<hostfr::fr*>(ctx->pinned);
    ZKP_CUDA(ctx, cudaMemcpyAsync(hp, S, sizeof(fr_t), cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(ctx, cudaStreamSynchronize(st));
    const hostfr::fr hi = hostfr::inv(*hp);
    fr_t inv_v;
    memcpy(inv_v.l, hi.l, 32);
    (void)inv;
    perm_z_combine_kernel<<<blocks_for(n, 128), 128, 0, st>>>(P, S, inv_v, n, out->d + out_off);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

int zkp_quotient_dev(zkp_ctx* ctx, unsigned k8, const zkp_quotient_args* q, zkp_buf* out, size_t out_off) {
    if (k8 > 28) return ZKP_ERR_INVALID;
    return zkp_quotient_range_dev(ctx, k8, q, 0, (size_t)1 << k8, out, out_off);
}

int zkp_quotient_range_dev(zkp_ctx* ctx, unsigned k8, const zkp_quotient_args* q, size_t first, size_t count,
                           zkp_buf* out, size_t out_off) {
    if (!ctx || !q || !out) return ZKP_ERR_INVALID;
    const bool cosets = q->coset_log_n != 0;
    if (cosets) {   // every vector holds `count` local points: whole cosets of n points
        if (q->coset_log_n > 25 || first != 0 || q->sliced || (count & (((size_t)1 << q->coset_log_n) - 1)) ||
            q->coset_first + (count >> q->coset_log_n) > 8)
            return ZKP_ERR_INVALID;
    } else if (k8 < 3 || k8 > 28) {
        return ZKP_ERR_INVALID;
    }
    const size_t n8 = cosets ? count : (size_t)1 << k8;
    if (out_off + n8 > out->n || first + count > n8) return ZKP_ERR_INVALID;
    QuotArgs a;
    auto ptr = [&](const zkp_poly_ref& r, const fr_t** p) {
        if (!CHECK_REF(r) || r.len < n8) return false;
        *p = r.buf->d + r.off;
        return true;
    };
    // witness-side vectors: whole, or (q->sliced) only the rank's slice plus an 8-element halo
    const bool sliced = q->sliced != 0;
    auto wptr = [&](const zkp_poly_ref& r, const fr_t** p) {
        if (!sliced) return ptr(r, p);
        if (!CHECK_REF(r) || r.len < count + 8) return false;
        *p = r.buf->d + r.off;
        return true;
    };
    bool ok = true;
    for (int j = 0; j < 4; j++) ok = ok && wptr(q->wires[j], &a.w[j]) && ptr(q->sigma[j], &a.sigma[j]);
    ok = ok && wptr(q->z, &a.z) && wptr(q->pi, &a.pi) && wptr(q->l1, &a.l1) && ptr(q->linear, &a.linear);
    for (int j = 0; j < 11; j++) ok = ok && ptr(q->sel[j], &a.sel[j]);
    if (!ok) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    a.alpha = fr_from_host(q->challenges[0]);
    a.beta = fr_from_host(q->challenges[1]);
    a.gamma = fr_from_host(q->challenges[2]);
    a.rs = fr_from_host(q->challenges[3]);
    a.ls = fr_from_host(q->challenges[4]);
    a.fs = fr_from_host(q->challenges[5]);
    a.vs = fr_from_host(q->challenges[6]);
    a.bk1 = a.beta * from_u64<FrParams>(7);
    a.bk2 = a.beta * from_u64<FrParams>(13);
    a.bk3 = a.beta * from_u64<FrParams>(17);
    // JubJub d = -(10240 / 10241)
    static const fr_t edwards_d = neg(from_u64<FrParams>(10240) * inverse(from_u64<FrParams>(10241)));  // once
    a.edwards_d = edwards_d;
    for (int j = 0; j < 8; j++) a.zh_inv[j] = fr_from_host(q->zh_inv[j]);
    a.mask = q->widget_mask;
    a.n8 = n8;
    a.first = first;
    a.count = count;
    a.sliced = sliced ? 1 : 0;
    a.coset_k = q->coset_log_n;
    a.coset_first = q->coset_first;
    a.out = out->d + out_off;
    if (count == 0) return ZKP_OK;
    ProfScope prof(ctx, "quotient");
    quotient_kernel<<<blocks_for(count, 128), 128, 0, ctx->stream>>>(a);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

}  // extern "C"

// launches only: the count results stay on the device (*res_dev, valid until the context's next scratch use)
int zkp::poly_eval2_launch(zkp_ctx* ctx, const zkp_poly_ref* polys, const uint8_t* which, unsigned count,
                           const uint64_t points[8], fr_t** res_dev) {
    if (!ctx || !polys || !points || !res_dev || count == 0 || count > EV_MAX) return ZKP_ERR_INVALID;
    size_t maxlen = 0;
    for (unsigned i = 0; i < count; i++) {
        if (!CHECK_REF(polys[i]) || (which && which[i] > 1)) return ZKP_ERR_INVALID;
        if (polys[i].len > maxlen) maxlen = polys[i].len;
    }
    int rc;
    if ((rc = set_device(ctx))) return rc;
    EvalArgs a;
    memset(&a, 0, sizeof a);
    bool second = false;
    for (unsigned i = 0; i < count; i++) {
        a.p[i] = polys[i].buf->d + polys[i].off;
        a.len[i] = polys[i].len;
        a.which[i] = which ? which[i] : 0;
        second = second || a.which[i];
    }
    for (int k = 0; k < (second ? 2 : 1); k++) {
        // 64-bit-limb host arithmetic for the 27 squarings (the device header's host emulation is 10x slower)
        hostfr::fr v;
        memcpy(v.l, points + 4 * k, 32);
        for (int j = 0; j < 28; j++) {
            memcpy(a.pw[k][j].l, v.l, 32);
            v = hostfr::mul(v, v);
        }
    }
    const size_t per_block = (size_t)EV_T * EV_L;
    a.nblocks = (unsigned)((maxlen + per_block - 1) / per_block);
    if (a.nblocks == 0) a.nblocks = 1;
    if ((size_t)a.nblocks > ((size_t)1 << 15)) return ZKP_ERR_INVALID;  // 2^28 coefficients: X^(2^27) is the last table entry
    fr_t* s;
    if ((rc = scratch(ctx, (size_t)count * a.nblocks + count, &s))) return rc;
    a.partial = s;
    fr_t* res = s + (size_t)count * a.nblocks;
    ProfScope prof(ctx, "poly_eval");
    eval_block_kernel<<<dim3(a.nblocks, count), EV_T, 0, ctx->stream>>>(a);
    ZKP_LAUNCHED(ctx);
    eval_final_kernel<<<count, EV_T, 0, ctx->stream>>>(a, res);
    ZKP_LAUNCHED(ctx);
    *res_dev = res;
    return ZKP_OK;
}

extern "C" {

int zkp_poly_eval2_dev(zkp_ctx* ctx, const zkp_poly_ref* polys, const uint8_t* which, unsigned count,
                       const uint64_t points[8], uint64_t* out) {
    if (!out) return ZKP_ERR_INVALID;
    fr_t* res = nullptr;
    int rc = poly_eval2_launch(ctx, polys, which, count, points, &res);
    if (rc) return rc;
    fr_t* h = reinterpret_cast<fr_t*>(ctx->pinned);
    ZKP_CUDA(ctx, cudaMemcpyAsync(h, res, count * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, h, count * sizeof(fr_t));
    return ZKP_OK;
}

int zkp_poly_eval_dev(zkp_ctx* ctx, const zkp_poly_ref* polys, unsigned count, const uint64_t point[4],
                      uint64_t* out) {
    if (!point) return ZKP_ERR_INVALID;
    uint64_t pts[8] = {point[0], point[1], point[2], point[3], 0, 0, 0, 0};
    return zkp_poly_eval2_dev(ctx, polys, nullptr, count, pts, out);
}

int zkp_poly_lincomb_dev(zkp_ctx* ctx, const zkp_poly_ref* polys, const uint64_t* scalars, unsigned count,
                         zkp_buf* out, size_t out_off, size_t out_len) {
    if (!ctx || !polys || !scalars || !out || count == 0 || count > LC_MAX || out_off + out_len > out->n)
        return ZKP_ERR_INVALID;
    LincombArgs a;
    memset(&a, 0, sizeof a);
    for (unsigned i = 0; i < count; i++) {
        if (!CHECK_REF(polys[i])) return ZKP_ERR_INVALID;
        a.p[i] = polys[i].buf->d + polys[i].off;
        a.len[i] = polys[i].len;
        a.s[i] = fr_from_host(scalars + 4 * i);
    }
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (!out_len) return ZKP_OK;
    a.count = count; a.n = out_len; a.out = out->d + out_off;
    ProfScope prof(ctx, "poly_lincomb");
    lincomb_kernel<<<blocks_for(out_len, 256), 256, 0, ctx->stream>>>(a);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

int zkp_poly_div_linear_dev(zkp_ctx* ctx, zkp_poly_ref in, const uint64_t point[4], zkp_buf* out, size_t out_off) {
    if (!ctx || !point || !out || !CHECK_REF(in) || in.len < 1 || out_off + in.len - 1 > out->n)
        return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    const size_t n = in.len;
    if (n == 1) return ZKP_OK;
    const size_t nc = (n + DV_L - 1) / DV_L;
    fr_t* s;
    if ((rc = scratch(ctx, nc + (nc + DV_TILE - 1) / DV_TILE, &s))) return rc;
    const fr_t pt = fr_from_host(point);
    const fr_t X = pow_u64(pt, DV_L);
    const fr_t* c = in.buf->d + in.off;
    cudaStream_t st = ctx->stream;
    ProfScope prof(ctx, "poly_div");
    div_chunk_kernel<<<blocks_for(nc, 128), 128, 0, st>>>(c, n, pt, s);
    ZKP_LAUNCHED(ctx);
    if (nc <= DV_TILE) {
        div_carry_kernel<<<1, 256, 0, st>>>(s, nc, nc, X, nullptr);
        ZKP_LAUNCHED(ctx);
        div_apply_kernel<<<blocks_for(nc, 128), 128, 0, st>>>(c, n, pt, s, nullptr, nc, X, out->d + out_off);
        ZKP_LAUNCHED(ctx);
    } else {
        const size_t ntiles = (nc + DV_TILE - 1) / DV_TILE;
        fr_t* tv = s + nc;
        div_carry_kernel<<<(unsigned)ntiles, 256, 0, st>>>(s, nc, DV_TILE, X, tv);
        ZKP_LAUNCHED(ctx);
        div_carry_kernel<<<1, ntiles <= 256 ? 32 : 256, 0, st>>>(tv, ntiles, ntiles, pow_u64(X, DV_TILE), nullptr);
        ZKP_LAUNCHED(ctx);
        div_apply_kernel<<<blocks_for(nc, 128), 128, 0, st>>>(c, n, pt, s, tv, nc, X, out->d + out_off);
        ZKP_LAUNCHED(ctx);
    }
    return ZKP_OK;
}

}  // extern "C"

// Radix-2 NTT / iNTT / coset NTT over BLS12-381 Fr for sm_100a.
//
// Replaces poly_commit::Fft::{new,dft,idft,coset_dft,coset_idft} as the reference calls
// them (src/prover.rs:87-88,121-124,192,229; src/prover/quotient_poly.rs:50-58,115,145,237;
// src/key.rs:83,121-131,222-245; src/permutation.rs:194-197,229-235): by-value transforms,
// short inputs zero-padded, natural-order output, coset shift g = 7, n^-1 on the inverse.
//
// Design (not a translation of the CPU recursion):
//  * log2(N) radix-2 decimation-in-frequency stages are grouped into P passes over HBM.
//    Pass i owns a "digit" of s_i index bits; a thread block stages a tile of
//    2^s_i (digit) x 2^c (neighbouring columns, for 128-byte coalescing) elements in
//    shared memory, runs s_i butterfly stages there, and writes the tile back.
//  * Every butterfly costs exactly one Montgomery multiplication: the twiddle of stage t
//    is read from a per-domain table w^j (j < N/2) precomputed in HBM, so there are no
//    separate inter-pass twiddle multiplies (the classical four-step penalty).
//  * Passes 1..P-1 are index-preserving (tile in, tile out); the last pass scatters each
//    tile to the digit-reversed address, which makes the overall transform
//    natural-order -> natural-order with no bit-reversal pass.  The scatter stays
//    128-byte coalesced because the tile's column bits are the low bits of the *first*
//    digit, i.e. the low bits of the output index.
//  * iNTT uses the same table: w^-j = -w^(N/2-j), folded into the butterfly as (v-u).
//  * Zero padding is applied on load; coset scaling (g^i on input of the forward
//    transform, g^-i n^-1 on output of the inverse) is fused into the first / last pass
//    through two-level power tables (g^i = lo[i & 1023] * hi[i >> 10]).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace zkp {

static constexpr unsigned S_MAX = 8;    // digit bits per pass
static constexpr unsigned C_LOG = 2;    // 4 columns x 32 B = 128 B segments
static constexpr unsigned LOG_LO = 10;  // low table of the two-level coset powers

struct NttDomain {
    unsigned k = 0;
    fr_t* tw = nullptr;     // w^j, j < max(1, N/2)
    fr_t* g_lo = nullptr;   // g^i,            i < 2^min(k,10)
    fr_t* g_hi = nullptr;   // g^(i << 10),    i < 2^(k-10)       (k > 10)
    fr_t* gi_lo = nullptr;  // n^-1 g^-i,      i < 2^min(k,10)
    fr_t* gi_hi = nullptr;  // g^-(i << 10)
    fr_t n_inv;
};

// ------------------------------------------------------------------ host-side constants
static fr_t host_root_of_unity() {
    // 7^((r-1)/2^32), canonical limbs (zkstd FftField::ROOT_OF_UNITY)
    fr_t raw;
    const uint32_t t[8] = {0x439f0d2bu, 0x3829971fu, 0x8c2280b9u, 0xb6368350u,
                           0x22c813b4u, 0xd09b6819u, 0xdfe81f20u, 0x16a2a19eu};
    for (int i = 0; i < 8; i++) raw.l[i] = t[i];
    return to_mont(raw);
}

fr_t fft_constant_host(unsigned k, int kind) {
    switch (kind) {
        case 0:
        case 1: {
            fr_t w = host_root_of_unity();
            for (unsigned i = k; i < 32; i++) w = sqr(w);
            return kind == 0 ? w : inverse(w);
        }
        case 2: return inverse(from_u64<FrParams>(1ull << k));
        case 3: return from_u64<FrParams>(7);
        default: return inverse(from_u64<FrParams>(7));
    }
}

// out[i] = first * base^i
__global__ void geometric_table_kernel(fr_t* out, size_t n, fr_t base, fr_t first) {
    constexpr unsigned B = 16;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * B;
    if (lo >= n) return;
    fr_t v = first * pow_u64(base, lo);
    for (unsigned i = 0; i < B && lo + i < n; i++) {
        out[lo + i] = v;
        v = v * base;
    }
}

static int build_table(zkp_ctx* ctx, fr_t** out, size_t n, const fr_t& base, const fr_t& first) {
    ZKP_CUDA(ctx, cudaMalloc(out, n * sizeof(fr_t)));
    size_t threads = (n + 15) / 16;
    unsigned block = 128;
    geometric_table_kernel<<<(unsigned)((threads + block - 1) / block), block, 0, ctx->stream>>>(
        *out, n, base, first);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

static int get_domain(zkp_ctx* ctx, unsigned k, NttDomain** out) {
    auto it = ctx->domains.find(k);
    if (it != ctx->domains.end()) { *out = it->second; return ZKP_OK; }
    NttDomain* d = new NttDomain();
    d->k = k;
    const size_t n = (size_t)1 << k;
    fr_t w = fft_constant_host(k, 0);
    fr_t g = fft_constant_host(k, 3), gi = fft_constant_host(k, 4);
    d->n_inv = fft_constant_host(k, 2);
    int rc;
    if ((rc = build_table(ctx, &d->tw, n > 1 ? n / 2 : 1, w, fr_t::one()))) return rc;
    const size_t nlo = (size_t)1 << (k < LOG_LO ? k : LOG_LO);
    if ((rc = build_table(ctx, &d->g_lo, nlo, g, fr_t::one()))) return rc;
    if ((rc = build_table(ctx, &d->gi_lo, nlo, gi, d->n_inv))) return rc;
    if (k > LOG_LO) {
        const size_t nhi = (size_t)1 << (k - LOG_LO);
        if ((rc = build_table(ctx, &d->g_hi, nhi, pow_u64(g, 1ull << LOG_LO), fr_t::one()))) return rc;
        if ((rc = build_table(ctx, &d->gi_hi, nhi, pow_u64(gi, 1ull << LOG_LO), fr_t::one()))) return rc;
    }
    ctx->domains[k] = d;
    *out = d;
    return ZKP_OK;
}

// ------------------------------------------------------------------ cosets of the n-domain inside the 8n domain
// The 8n-point quotient domain g <w_8n> is the union of the eight cosets h_u H_n, h_u = g w_8n^u, and
// point 8 m + u of the reference's ordering is h_u w_n^m.  A polynomial's values on one coset are an
// n-point transform of its coefficients scaled by h_u^e (coefficients beyond n fold back with h_u^n), and
// the 8n-point inverse is eight n-point inverses, a scaling by h_u^-e / 8 and an 8 x 8 combination across
// cosets (csrc/create_proof.cu) -- which is what lets a GPU own whole cosets of the quotient.
__device__ __forceinline__ fr_t ld_fr(const fr_t* p);
__device__ __forceinline__ fr_t ldg_fr(const fr_t* p);
__device__ __forceinline__ void st_fr(fr_t* p, const fr_t& v);

struct TwTab { fr_t* lo = nullptr; fr_t* hi = nullptr; unsigned h = 0; };   // four-step twiddles (below)

struct Coset8Tab {
    fr_t* f_lo = nullptr;   // h^e,           e < 2^min(k,10)
    fr_t* f_hi = nullptr;   // h^(e << 10)
    fr_t* b_lo = nullptr;   // h^-e / 8
    fr_t* b_hi = nullptr;   // h^-(e << 10)
    fr_t fold;              // h^n
};

static int get_coset8(zkp_ctx* ctx, unsigned k, unsigned u, Coset8Tab** out) {
    const unsigned key = k * 8 + u;
    auto it = ctx->coset8.find(key);
    if (it != ctx->coset8.end()) { *out = it->second; return ZKP_OK; }
    Coset8Tab* t = new Coset8Tab();
    const fr_t h = fft_constant_host(k, 3) * pow_u64(fft_constant_host(k + 3, 0), u);
    const fr_t hi = inverse(h);
    t->fold = pow_u64(h, 1ull << k);
    const size_t nlo = (size_t)1 << (k < LOG_LO ? k : LOG_LO);
    int rc;
    if ((rc = build_table(ctx, &t->f_lo, nlo, h, fr_t::one()))) return rc;
    if ((rc = build_table(ctx, &t->b_lo, nlo, hi, inverse(from_u64<FrParams>(8))))) return rc;
    if (k > LOG_LO) {
        const size_t nhi = (size_t)1 << (k - LOG_LO);
        if ((rc = build_table(ctx, &t->f_hi, nhi, pow_u64(h, 1ull << LOG_LO), fr_t::one()))) return rc;
        if ((rc = build_table(ctx, &t->b_hi, nhi, pow_u64(hi, 1ull << LOG_LO), fr_t::one()))) return rc;
    }
    ctx->coset8[key] = t;
    *out = t;
    return ZKP_OK;
}

struct Coset8Args {
    const fr_t* lo[8];
    const fr_t* hi[8];
    fr_t fold[8];
    const fr_t* in;
    fr_t* out;
    size_t in_stride, len_in, n;
    unsigned k;
    int use_fold;
};

// out[j][e] = (in_j[e] + fold_j in_j[n + e]) * tab_j(e), e < n; in_j = in + j in_stride (stride 0: one input)
__global__ void __launch_bounds__(256) coset8_scale_kernel(const __grid_constant__ Coset8Args a) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned j = blockIdx.y;
    if (e >= a.n) return;
    const fr_t* in = a.in + (size_t)j * a.in_stride;
    fr_t v = e < a.len_in ? ld_fr(in + e) : fr_t::zero();
    if (a.use_fold && a.n + e < a.len_in) v = v + a.fold[j] * ld_fr(in + a.n + e);
    fr_t f = ldg_fr(a.lo[j] + (e & ((1u << LOG_LO) - 1)));
    if (a.k > LOG_LO) f = f * ldg_fr(a.hi[j] + (e >> LOG_LO));
    st_fr(a.out + (size_t)j * a.n + e, v * f);
}

// values of one polynomial (len_in <= 2n coefficients) on the cosets first .. first + count - 1: out[j][m]
int coset8_forward(zkp_ctx* ctx, const fr_t* in, size_t len_in, fr_t* out, unsigned k, unsigned first, unsigned count) {
    const size_t n = (size_t)1 << k;
    if (k > 25 || count == 0 || first + count > 8 || len_in > 2 * n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    Coset8Args a;
    memset(&a, 0, sizeof a);
    for (unsigned j = 0; j < count; j++) {
        Coset8Tab* t;
        if ((rc = get_coset8(ctx, k, first + j, &t))) return rc;
        a.lo[j] = t->f_lo; a.hi[j] = t->f_hi; a.fold[j] = t->fold;
    }
    a.in = in; a.out = out; a.in_stride = 0; a.len_in = len_in; a.n = n; a.k = k; a.use_fold = 1;
    {
        ProfScope prof(ctx, "ntt");
        coset8_scale_kernel<<<dim3((unsigned)((n + 255) / 256), count), 256, 0, ctx->stream>>>(a);
        ZKP_LAUNCHED(ctx);
    }
    return ntt_run(ctx, out, n, n, out, n, k, false, false, count);
}

// data[j][.] (values on coset first + j) -> inverse transform of each coset scaled by h_u^-e / 8, in place:
// the per-coset term of the 8n-point inverse, ready for the cross-coset combination
int coset8_inverse_local(zkp_ctx* ctx, fr_t* data, unsigned k, unsigned first, unsigned count) {
    const size_t n = (size_t)1 << k;
    if (k > 25 || count == 0 || first + count > 8) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = ntt_run(ctx, data, n, n, data, n, k, true, false, count))) return rc;
    Coset8Args a;
    memset(&a, 0, sizeof a);
    for (unsigned j = 0; j < count; j++) {
        Coset8Tab* t;
        if ((rc = get_coset8(ctx, k, first + j, &t))) return rc;
        a.lo[j] = t->b_lo; a.hi[j] = t->b_hi;
    }
    a.in = data; a.out = data; a.in_stride = n; a.len_in = n; a.n = n; a.k = k; a.use_fold = 0;
    ProfScope prof(ctx, "ntt");
    coset8_scale_kernel<<<dim3((unsigned)((n + 255) / 256), count), 256, 0, ctx->stream>>>(a);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

void ntt_free_domains(zkp_ctx* ctx) {
    for (auto& kv : ctx->coset8) {
        Coset8Tab* t = kv.second;
        cudaFree(t->f_lo); cudaFree(t->f_hi); cudaFree(t->b_lo); cudaFree(t->b_hi);
        delete t;
    }
    ctx->coset8.clear();
    for (auto& kv : ctx->twtabs) { cudaFree(kv.second->lo); cudaFree(kv.second->hi); delete kv.second; }
    ctx->twtabs.clear();
    for (auto& kv : ctx->domains) {
        NttDomain* d = kv.second;
        cudaFree(d->tw); cudaFree(d->g_lo); cudaFree(d->g_hi); cudaFree(d->gi_lo); cudaFree(d->gi_hi);
        delete d;
    }
    ctx->domains.clear();
    if (ctx->ntt_scratch) cudaFree(ctx->ntt_scratch);
    ctx->ntt_scratch = nullptr;
    ctx->ntt_scratch_n = 0;
}

// ------------------------------------------------------------------ the pass kernel
struct PassParams {
    const fr_t* in;
    fr_t* out;
    size_t in_stride, out_stride, len_in;
    const fr_t* tw;
    const fr_t* sc_lo;
    const fr_t* sc_hi;
    fr_t n_inv;
    unsigned k, lo, s, c_log, cpos, tw_shift;
    unsigned ndig, dig[6];
    int last, inverse;
    int scale_in;   // forward coset: a_i *= g^i on load
    int scale_out;  // 0 none, 1 constant n^-1, 2 table n^-1 g^-i
};

__device__ __forceinline__ fr_t ld_fr(const fr_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ fr_t ldg_fr(const fr_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(fr_t* p, const fr_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// shared-memory tile: low and high 16-byte halves in separate planes (conflict-free
// 128-bit accesses for unit-stride element indices)
__device__ __forceinline__ fr_t lds_fr(const uint4* lo, const uint4* hi, unsigned e) {
    uint4 a = lo[e], b = hi[e];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fr(uint4* lo, uint4* hi, unsigned e, const fr_t& v) {
    lo[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

__device__ __forceinline__ size_t insert_bits(size_t x, unsigned pos, unsigned len, size_t val) {
    return ((x >> pos) << (pos + len)) | (val << pos) | (x & (((size_t)1 << pos) - 1));
}

__device__ __forceinline__ size_t tile_index(const PassParams& p, size_t rest, unsigned d, unsigned c) {
    if (p.last) {  // digit at bit 0, columns at cpos >= s
        size_t x = insert_bits(rest, 0, p.s, d);
        return insert_bits(x, p.cpos, p.c_log, c);
    }
    size_t x = insert_bits(rest, 0, p.c_log, c);
    return insert_bits(x, p.lo, p.s, d);
}

__device__ __forceinline__ size_t digit_reverse(const PassParams& p, size_t idx) {
    size_t o = 0;
    unsigned pos = p.k, sh = 0;
    for (unsigned i = 0; i < p.ndig; i++) {
        pos -= p.dig[i];
        o |= ((idx >> pos) & (((size_t)1 << p.dig[i]) - 1)) << sh;
        sh += p.dig[i];
    }
    return o;
}

__device__ __forceinline__ fr_t scale_factor(const PassParams& p, size_t i) {
    fr_t f = ldg_fr(p.sc_lo + (i & ((1u << LOG_LO) - 1)));
    if (p.k > LOG_LO) f = f * ldg_fr(p.sc_hi + (i >> LOG_LO));
    return f;
}

// One radix-2 DIF butterfly on the pair (u, v) with twiddle index widx (w^widx; the inverse
// transform uses w^-widx = -w^(N/2 - widx)): u <- u + v, v <- (u - v) w^(+-widx).
__device__ __forceinline__ void dif_butterfly(fr_t& u, fr_t& v, size_t widx, const PassParams& p, size_t half) {
    const fr_t sum = u + v;
    if (widx == 0) {
        v = u - v;
    } else if (p.inverse) {
        v = (v - u) * ldg_fr(p.tw + (half - widx));
    } else {
        v = (u - v) * ldg_fr(p.tw + widx);
    }
    u = sum;
}

// blockDim = T / 4 threads for a tile of T = 2^(s + c_log) elements: every thread keeps four
// elements in registers and runs TWO butterfly stages on them (a radix-4 step: 4 multiplications,
// 3 distinct twiddles) between shared-memory exchanges, halving the shared-memory traffic and
// the barriers of a stage-per-barrier radix-2 loop; an odd stage count ends with one radix-2 step.
// MB bounds the register allocation for MB resident blocks per SM (3: 80 registers, no spill).
template <int MB>
__global__ void __launch_bounds__(256, MB) ntt_pass_kernel(PassParams p) {
    extern __shared__ uint4 smem[];
    const unsigned T = 1u << (p.s + p.c_log);
    uint4* s_lo = smem;
    uint4* s_hi = smem + T;
    const unsigned tid = threadIdx.x;
    const size_t rest = blockIdx.x;
    const fr_t* in = p.in + (size_t)blockIdx.y * p.in_stride;
    fr_t* out = p.out + (size_t)blockIdx.y * p.out_stride;
    const unsigned C = 1u << p.c_log;

    for (unsigned e = tid; e < T; e += blockDim.x) {
        const unsigned c = e & (C - 1), d = e >> p.c_log;
        const size_t idx = tile_index(p, rest, d, c);
        fr_t v = fr_t::zero();
        if (idx < p.len_in) {
            v = ld_fr(in + idx);
            if (p.scale_in) v = v * scale_factor(p, idx);
        }
        sts_fr(s_lo, s_hi, e, v);
    }
    __syncthreads();

    {
        // index bits below the digit (zero on the last pass)
        const size_t half = ((size_t)1 << p.k) >> 1;
        unsigned t = 0;
        for (; t + 1 < p.s; t += 2) {  // radix-4 step: stages t and t + 1
            const unsigned span = p.s - 1 - t;  // >= 1
            for (unsigned g = tid; g < (T >> 2); g += blockDim.x) {
                const unsigned c = g & (C - 1), q = g >> p.c_log;
                const size_t low = p.last ? 0 : ((rest & (((size_t)1 << (p.lo - p.c_log)) - 1)) << p.c_log) | c;
                const unsigned j0 = q & ((1u << (span - 1)) - 1);
                const unsigned d0 = ((q >> (span - 1)) << (span + 1)) | j0;
                const unsigned hstep = 1u << (span - 1 + p.c_log);
                const unsigned e0 = (d0 << p.c_log) | c, e1 = e0 + hstep, e2 = e0 + 2 * hstep, e3 = e2 + hstep;
                fr_t x0 = lds_fr(s_lo, s_hi, e0), x1 = lds_fr(s_lo, s_hi, e1);
                fr_t x2 = lds_fr(s_lo, s_hi, e2), x3 = lds_fr(s_lo, s_hi, e3);
                const size_t ja = ((size_t)j0 << p.lo) | low;
                const size_t jb = ((size_t)(j0 + (1u << (span - 1))) << p.lo) | low;
                // stage t: (x0, x2) and (x1, x3)
                dif_butterfly(x0, x2, (ja << t) << p.tw_shift, p, half);
                dif_butterfly(x1, x3, (jb << t) << p.tw_shift, p, half);
                // stage t + 1: (x0, x1) and (x2, x3); the very last stage of the transform has w = 1
                const size_t w1 = (p.lo == 0 && span == 1) ? 0 : ((ja << (t + 1)) << p.tw_shift);
                dif_butterfly(x0, x1, w1, p, half);
                dif_butterfly(x2, x3, w1, p, half);
                sts_fr(s_lo, s_hi, e0, x0); sts_fr(s_lo, s_hi, e1, x1);
                sts_fr(s_lo, s_hi, e2, x2); sts_fr(s_lo, s_hi, e3, x3);
            }
            __syncthreads();
        }
        if (t < p.s) {  // one radix-2 stage left (span = 0)
            for (unsigned b = tid; b < (T >> 1); b += blockDim.x) {
                const unsigned c = b & (C - 1), q = b >> p.c_log;
                const size_t low = p.last ? 0 : ((rest & (((size_t)1 << (p.lo - p.c_log)) - 1)) << p.c_log) | c;
                const unsigned e0 = ((q << 1) << p.c_log) | c, e1 = e0 + C;
                fr_t u = lds_fr(s_lo, s_hi, e0), v = lds_fr(s_lo, s_hi, e1);
                const size_t widx = p.lo == 0 ? 0 : ((low << t) << p.tw_shift);
                dif_butterfly(u, v, widx, p, half);
                sts_fr(s_lo, s_hi, e0, u);
                sts_fr(s_lo, s_hi, e1, v);
            }
            __syncthreads();
        }
    }

    for (unsigned e = tid; e < T; e += blockDim.x) {
        const unsigned c = e & (C - 1), d = e >> p.c_log;
        const unsigned kd = p.s ? (__brev(d) >> (32 - p.s)) : 0;  // DIF leaves the digit bit-reversed
        size_t idx = tile_index(p, rest, kd, c);
        fr_t v = lds_fr(s_lo, s_hi, e);
        if (p.last) {
            idx = digit_reverse(p, idx);
            if (p.scale_out == 1) v = v * p.n_inv;
            else if (p.scale_out == 2) v = v * scale_factor(p, idx);
        }
        st_fr(out + idx, v);
    }
}

// ------------------------------------------------------------------ planning + launch
int ntt_run(zkp_ctx* ctx, const fr_t* in, size_t in_stride, size_t len_in, fr_t* out,
            size_t out_stride, unsigned k, bool inverse, bool coset, unsigned batch) {
    if (k > 28 || batch == 0) return ZKP_ERR_INVALID;
    const size_t n = (size_t)1 << k;
    if (len_in > n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    NttDomain* dom;
    if ((rc = get_domain(ctx, k, &dom))) return rc;

    const unsigned P = k == 0 ? 1 : (k + S_MAX - 1) / S_MAX;
    unsigned dig[6];
    for (unsigned i = 0; i < P; i++) dig[i] = k / P + (i < k % P ? 1 : 0);

    const fr_t* src = in;
    size_t src_stride = in_stride;
    fr_t* work = nullptr;
    if (P > 1) {
        const size_t need = n * batch;
        if (ctx->ntt_scratch_n < need) {
            if (ctx->ntt_scratch) cudaFree(ctx->ntt_scratch);
            ctx->ntt_scratch = nullptr; ctx->ntt_scratch_n = 0;
            ZKP_CUDA(ctx, cudaMalloc(&ctx->ntt_scratch, need * sizeof(fr_t)));
            ctx->ntt_scratch_n = need;
        }
        work = ctx->ntt_scratch;
    }

    ProfScope prof(ctx, "ntt");
    unsigned lo = k;
    for (unsigned i = 0; i < P; i++) {
        lo -= dig[i];
        PassParams p;
        p.k = k; p.s = dig[i]; p.lo = lo;
        p.last = (i == P - 1);
        p.inverse = inverse;
        p.tw = dom->tw;
        p.tw_shift = k - p.s - lo;
        p.ndig = P;
        for (unsigned j = 0; j < 6; j++) p.dig[j] = j < P ? dig[j] : 0;
        if (p.last) {
            p.c_log = P > 1 ? (dig[0] < C_LOG ? dig[0] : C_LOG) : 0;
            p.cpos = k - dig[0];
        } else {
            p.c_log = lo < C_LOG ? lo : C_LOG;
            p.cpos = 0;
        }
        p.in = src; p.in_stride = src_stride;
        p.len_in = (i == 0) ? len_in : n;
        if (p.last) { p.out = out; p.out_stride = out_stride; }
        else { p.out = work; p.out_stride = n; }
        p.scale_in = (i == 0 && coset && !inverse);
        p.scale_out = 0;
        p.sc_lo = nullptr; p.sc_hi = nullptr;
        p.n_inv = dom->n_inv;
        if (p.scale_in) { p.sc_lo = dom->g_lo; p.sc_hi = dom->g_hi; }
        if (p.last && inverse) {
            p.scale_out = coset ? 2 : 1;
            if (coset) { p.sc_lo = dom->gi_lo; p.sc_hi = dom->gi_hi; }
        }
        const unsigned tlog = p.s + p.c_log;
        const unsigned threads = tlog >= 2 ? (1u << (tlog - 2)) : 1;
        const size_t smem = ((size_t)2 << tlog) * sizeof(uint4);
        dim3 grid((unsigned)(n >> tlog), batch);
        static int mb = 0;
        if (!mb) {
            mb = 3;
            if (const char* e = getenv("ZKP_NTT_BLOCKS_PER_SM")) mb = atoi(e);  // tuning knob: 2..4
            if (mb < 2) mb = 2;
            if (mb > 4) mb = 4;
        }
        const unsigned nthreads = threads < 32 ? 32 : threads;
        if (mb == 2) ntt_pass_kernel<2><<<grid, nthreads, smem, ctx->stream>>>(p);
        else if (mb == 3) ntt_pass_kernel<3><<<grid, nthreads, smem, ctx->stream>>>(p);
        else ntt_pass_kernel<4><<<grid, nthreads, smem, ctx->stream>>>(p);
        ZKP_LAUNCHED(ctx);
        src = p.out; src_stride = p.out_stride;
    }
    return ZKP_OK;
}

// ------------------------------------------------------------------ four-step helpers (multi-GPU)
// out[a][b][0..w) = in[b][a][0..w): transpose of a B x A matrix of w-element blocks.  w == 1 goes
// through a shared-memory tile so both sides stay coalesced.
__global__ void __launch_bounds__(256) permute_blocks_kernel(const fr_t* in, fr_t* out, size_t A, size_t B, size_t w) {
    const size_t total = A * B * w;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t e = idx % w, ab = idx / w;
        const size_t b = ab % B, a = ab / B;  // idx enumerates out[a][b][e]
        st_fr(out + idx, ld_fr(in + (b * A + a) * w + e));
    }
}

__global__ void __launch_bounds__(256) transpose_tile_kernel(const fr_t* in, fr_t* out, size_t A, size_t B) {
    // in: B x A row-major, out: A x B row-major; 16 x 16 tiles
    __shared__ uint4 t_lo[16][17], t_hi[16][17];
    const unsigned tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t a0 = (size_t)blockIdx.x * 16, b0 = (size_t)blockIdx.y * 16;
    if (b0 + ty < B && a0 + tx < A) {
        const uint4* q = reinterpret_cast<const uint4*>(in + (b0 + ty) * A + a0 + tx);
        t_lo[ty][tx] = q[0];
        t_hi[ty][tx] = q[1];
    }
    __syncthreads();
    if (a0 + ty < A && b0 + tx < B) {
        uint4* q = reinterpret_cast<uint4*>(out + (a0 + ty) * B + b0 + tx);
        q[0] = t_lo[tx][ty];
        q[1] = t_hi[tx][ty];
    }
}

// data[a][b] *= f(a, b) over a rows x cols matrix, 16 consecutive b per thread.
//   mode 0 (four-step twiddle):  f = base1 ^ ((a0 + a) * b)
//   mode 1 (coset / scaling):    f = base1 ^ (a0 + a) * base2 ^ b
__global__ void __launch_bounds__(128) scale_matrix_kernel(fr_t* data, size_t rows, size_t cols, size_t a0, fr_t base1,
                                                          fr_t base2, int mode) {
    constexpr size_t RUN = 16;
    const size_t runs = (cols + RUN - 1) / RUN;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * runs) return;
    const size_t a = t / runs, b0 = (t % runs) * RUN;
    const fr_t pa = pow_u64(base1, (uint64_t)(a0 + a));
    const fr_t ratio = mode == 0 ? pa : base2;
    fr_t f = mode == 0 ? pow_u64(pa, (uint64_t)b0) : pa * pow_u64(base2, (uint64_t)b0);
    fr_t* row = data + a * cols;
    for (size_t b = b0; b < b0 + RUN && b < cols; b++) {
        st_fr(row + b, ld_fr(row + b) * f);
        f = f * ratio;
    }
}

// Four-step twiddle fused into the transpose: out[b][a] = in[a][b] * w_N^(+-(a0 + a) b), a < rows, b < cols.
// The exponent e = (a0 + a) b < N indexes a two-level table (w^e = lo[e & (2^h - 1)] hi[e >> h], h = ceil(k / 2)):
// two multiplications per element instead of the ~5 of per-thread powers, and one pass over the data
// instead of two (scale, then transpose).
static int get_twtab(zkp_ctx* ctx, unsigned k, bool inverse, TwTab** out) {
    const unsigned key = 64 * 8 + k * 2 + (inverse ? 1 : 0);     // shares the coset8 map's key space above 8 * 64
    auto it = ctx->twtabs.find(key);
    if (it != ctx->twtabs.end()) { *out = it->second; return ZKP_OK; }
    TwTab* t = new TwTab();
    t->h = (k + 1) / 2;
    const fr_t w = fft_constant_host(k, inverse ? 1 : 0);
    int rc;
    if ((rc = build_table(ctx, &t->lo, (size_t)1 << t->h, w, fr_t::one()))) return rc;
    if ((rc = build_table(ctx, &t->hi, (size_t)1 << (k - t->h), pow_u64(w, 1ull << t->h), fr_t::one()))) return rc;
    ctx->twtabs[key] = t;
    *out = t;
    return ZKP_OK;
}

__global__ void __launch_bounds__(256) twiddle_transpose_kernel(const fr_t* in, fr_t* out, size_t A, size_t B, size_t a0,
                                                               const fr_t* lo, const fr_t* hi, unsigned h) {
    // in: A x B row-major (a rows), out: B x A; 16 x 16 tiles through shared memory, both sides coalesced
    __shared__ uint4 t_lo[16][17], t_hi[16][17];
    const unsigned tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const size_t b0 = (size_t)blockIdx.x * 16, r0 = (size_t)blockIdx.y * 16;
    if (r0 + ty < A && b0 + tx < B) {
        const size_t a = r0 + ty, b = b0 + tx;
        fr_t v = ld_fr(in + a * B + b);
        const size_t e = (a0 + a) * b;
        v = v * (ldg_fr(lo + (e & (((size_t)1 << h) - 1))) * ldg_fr(hi + (e >> h)));
        t_lo[ty][tx] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        t_hi[ty][tx] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
    __syncthreads();
    if (b0 + ty < B && r0 + tx < A) {
        uint4* q = reinterpret_cast<uint4*>(out + (b0 + ty) * A + r0 + tx);
        q[0] = t_lo[tx][ty];
        q[1] = t_hi[tx][ty];
    }
}

int ntt_twiddle_transpose(zkp_ctx* ctx, const fr_t* in, fr_t* out, size_t rows, size_t cols, size_t a0, unsigned k,
                          bool inverse) {
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (rows == 0 || cols == 0) return ZKP_OK;
    if (k > 28 || (a0 + rows - 1) * (cols - 1) >= ((size_t)1 << k)) return ZKP_ERR_INVALID;
    TwTab* t;
    if ((rc = get_twtab(ctx, k, inverse, &t))) return rc;
    ProfScope prof(ctx, "ntt");
    dim3 grid((unsigned)((cols + 15) / 16), (unsigned)((rows + 15) / 16));
    twiddle_transpose_kernel<<<grid, 256, 0, ctx->stream>>>(in, out, rows, cols, a0, t->lo, t->hi, t->h);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

int ntt_permute(zkp_ctx* ctx, const fr_t* in, fr_t* out, size_t A, size_t B, size_t w) {
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (A == 0 || B == 0 || w == 0) return ZKP_OK;
    if (w == 1) {
        dim3 grid((unsigned)((A + 15) / 16), (unsigned)((B + 15) / 16));
        transpose_tile_kernel<<<grid, 256, 0, ctx->stream>>>(in, out, A, B);
    } else {
        const size_t total = A * B * w;
        size_t blocks = (total + 255) / 256;
        const size_t cap = (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 64;
        if (blocks > cap) blocks = cap;
        permute_blocks_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, out, A, B, w);
    }
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

int ntt_scale_matrix(zkp_ctx* ctx, fr_t* data, size_t rows, size_t cols, size_t a0, const fr_t& base1,
                     const fr_t& base2, int mode) {
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (rows == 0 || cols == 0) return ZKP_OK;
    const size_t threads = rows * ((cols + 15) / 16);
    scale_matrix_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(data, rows, cols, a0, base1, base2,
                                                                                   mode);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

int ntt_elements(zkp_ctx* ctx, unsigned k, fr_t* out) {
    int rc;
    if ((rc = set_device(ctx))) return rc;
    const size_t n = (size_t)1 << k;
    size_t threads = (n + 15) / 16;
    geometric_table_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(
        out, n, fft_constant_host(k, 0), fr_t::one());
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

}  // namespace zkp

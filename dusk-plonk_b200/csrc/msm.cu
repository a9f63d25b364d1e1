// G1 multi-scalar multiplication (Pippenger bucket method) for sm_100a.
//
// Replaces poly_commit::msm_curve_addition / PlonkParams::commit (call sites
// src/prover.rs:133-136,194,262-265,440,452; src/key.rs:138-159; src/prover/proof.rs:507-526):
// sum_i s_i * P_i over BLS12-381 G1, returned in affine form (Commitment::new).
//
// KZG commits always run against the same SRS, so the window structure is moved into the
// bases once, at load time (PlonkParams::trim): the device keeps the table
//     T[w][i] = 2^(c w) * P_i        (affine, W = floor(255 / c) + 1 copies)
// and every (scalar, window) pair becomes one signed digit d in [-2^(c-1), 2^(c-1)] that adds
// +-T[w][i] to bucket |d| - 1 of a SINGLE bucket set shared by all windows.  There is no
// per-window bucket array, no window-combine doubling chain, and the bucket reduction runs once.
//
// Pipeline (all on the device, one stream):
//   1. msm_digits_kernel      Fr Montgomery -> canonical, signed digits, bucket histogram
//   2. msm_scan_kernel        bucket offsets; buckets are cut into tasks of <= CAP entries so
//                             skewed digit distributions (top window, small coefficients,
//                             carry-only digits) cannot serialise on one thread
//   3. msm_scatter_kernel     counting sort of (table index, sign) by bucket
//   4. msm_accumulate_kernel  one thread per task, XYZZ mixed additions (8M + 2S each, 384-bit
//                             Montgomery on the integer pipe) -- the hot kernel
//   5. msm_merge_kernel       buckets made of several tasks: one warp sums the partials
//   6. msm_plane_sum / plane_combine / final_sum   sum_b (b+1) * B[b] as c bit-plane tree sums
//   7. host: one Fermat inversion -> affine
// Addition in G1 is commutative and the result is normalised to affine, so the output is
// bit-identical to any correct CPU evaluation regardless of accumulation order.
#include "common.cuh"

namespace zkp {

struct MsmScratch {
    size_t cap_entries = 0, cap_sorted = 0, cap_buckets = 0, cap_partials = 0, cap_tasks = 0;
    uint32_t* digits = nullptr;   // [W * n]  bucket | sign << 31, 0xffffffff = zero digit
    uint32_t* sorted = nullptr;   // [W * n]  table index | sign << 31, grouped by bucket
    uint32_t* counts = nullptr;   // [B]
    uint32_t* offsets = nullptr;  // [B]
    uint32_t* cursor = nullptr;   // [B]
    uint32_t* task_off = nullptr; // [B + 1]  first task of each bucket; [B] = number of tasks
    uint32_t* multi = nullptr;    // [B + 1]  buckets made of > 1 task; [B] = how many
    g1_xyzz* buckets = nullptr;   // [B]
    g1_xyzz* task_out = nullptr;  // [tasks] partial sums of multi-task buckets
    g1_xyzz* part_a = nullptr;    // reduction ping-pong
    g1_xyzz* part_b = nullptr;
    g1_affine* result = nullptr;
    long long* top = nullptr;
};

static constexpr uint32_t DIGIT_ZERO = 0xffffffffu;

__device__ __forceinline__ fr_t msm_ld_fr(const fr_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

__device__ __forceinline__ g1_affine msm_ld_affine(const g1_affine* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    g1_affine r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 6; i++) {
        uint4 v = __ldg(q + i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    return r;
}

// bits [pos, pos+c) of a 256-bit little-endian value (c <= 24)
__device__ __forceinline__ uint32_t window_bits(const fr_t& s, unsigned pos, unsigned c) {
    const unsigned limb = pos >> 5, off = pos & 31;
    if (limb >= 8) return 0;
    uint64_t v = s.l[limb];
    if (limb + 1 < 8) v |= (uint64_t)s.l[limb + 1] << 32;
    return (uint32_t)(v >> off) & ((1u << c) - 1);
}

// digits[w * n + i]: bucket | sign for scalar i, window w.
__global__ void msm_digits_kernel(const fr_t* scalars, size_t n, unsigned c, unsigned W,
                                  uint32_t* digits, uint32_t* counts) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr_t s = from_mont(msm_ld_fr(scalars + i));
    const uint32_t B = 1u << (c - 1);
    uint32_t carry = 0;
    for (unsigned w = 0; w < W; w++) {
        uint32_t d = window_bits(s, w * c, c) + carry;
        uint32_t enc;
        if (d > B) {  // represent as d - 2^c (negative or zero), carry into the next window
            const uint32_t m = (1u << c) - d;  // |d - 2^c| in [0, B-1]
            enc = m ? ((m - 1) | 0x80000000u) : DIGIT_ZERO;
            carry = 1;
        } else {
            carry = 0;
            enc = d ? (d - 1) : DIGIT_ZERO;
        }
        digits[(size_t)w * n + i] = enc;
        if (enc != DIGIT_ZERO) atomicAdd(&counts[enc & 0x7fffffffu], 1u);
    }
}

// A bucket of up to 2 * cap entries is one task; larger ones are cut into ceil(count / cap).
__host__ __device__ __forceinline__ uint32_t msm_ntasks(uint32_t count, uint32_t cap) {
    return count <= 2 * cap ? (count ? 1u : 0u) : (count + cap - 1) / cap;
}

// Single-block exclusive scan over the B bucket counts: entry offsets, task offsets
// (ceil(count / cap) tasks per bucket) and the list of buckets that need a merge.
__global__ void __launch_bounds__(1024) msm_scan_kernel(const uint32_t* counts, uint32_t* offsets, uint32_t* cursor,
                                                        uint32_t* task_off, uint32_t* multi, uint32_t B, uint32_t cap) {
    __shared__ uint32_t ws_e[32], ws_t[32];
    __shared__ uint32_t n_multi;
    const unsigned T = blockDim.x, tid = threadIdx.x;
    const uint32_t per = (B + T - 1) / T;
    const uint32_t lo = tid * per < B ? tid * per : B, hi = lo + per < B ? lo + per : B;
    uint32_t se = 0, stk = 0;
    for (uint32_t i = lo; i < hi; i++) { const uint32_t cn = counts[i]; se += cn; stk += msm_ntasks(cn, cap); }
    uint32_t ie = se, it = stk;
    const unsigned lane = tid & 31, wid = tid >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t ve = __shfl_up_sync(0xffffffffu, ie, o), vt = __shfl_up_sync(0xffffffffu, it, o);
        if (lane >= (unsigned)o) { ie += ve; it += vt; }
    }
    if (lane == 31) { ws_e[wid] = ie; ws_t[wid] = it; }
    if (tid == 0) n_multi = 0;
    __syncthreads();
    if (wid == 0) {
        uint32_t e = lane < (T >> 5) ? ws_e[lane] : 0, t = lane < (T >> 5) ? ws_t[lane] : 0;
        uint32_t xe = e, xt = t;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t ve = __shfl_up_sync(0xffffffffu, xe, o), vt = __shfl_up_sync(0xffffffffu, xt, o);
            if (lane >= (unsigned)o) { xe += ve; xt += vt; }
        }
        ws_e[lane] = xe - e;  // exclusive
        ws_t[lane] = xt - t;
        if (lane == 31) task_off[B] = xt;  // total number of tasks
    }
    __syncthreads();
    uint32_t re = ws_e[wid] + (ie - se), rt = ws_t[wid] + (it - stk);
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t cn = counts[i], nt = msm_ntasks(cn, cap);
        offsets[i] = re;
        cursor[i] = re;
        task_off[i] = rt;
        if (nt > 1) multi[atomicAdd(&n_multi, 1u)] = i;
        re += cn;
        rt += nt;
    }
    __syncthreads();
    if (tid == 0) multi[B] = n_multi;
}

// sorted[pos] = (w * stride + i) | sign, grouped by bucket
__global__ void msm_scatter_kernel(const uint32_t* digits, size_t n, unsigned W, uint32_t stride, uint32_t* cursor,
                                   uint32_t* sorted) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (unsigned w = 0; w < W; w++) {
        const uint32_t enc = digits[(size_t)w * n + i];
        if (enc == DIGIT_ZERO) continue;
        const uint32_t pos = atomicAdd(&cursor[enc & 0x7fffffffu], 1u);
        sorted[pos] = (w * stride + (uint32_t)i) | (enc & 0x80000000u);
    }
}

// One thread per task: sum of +-T[idx] over <= cap consecutive entries of one bucket.
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const g1_affine* table, const uint32_t* sorted,
                                                            const uint32_t* offsets, const uint32_t* counts,
                                                            const uint32_t* task_off, uint32_t B, uint32_t cap,
                                                            g1_xyzz* buckets, g1_xyzz* task_out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= task_off[B]) return;
    // owner bucket: task_off[b] <= t < task_off[b + 1]
    uint32_t lo = 0, hi = B;  // invariant: task_off[lo] <= t, task_off[hi] > t
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (task_off[mid] <= t) lo = mid; else hi = mid;
    }
    const uint32_t b = lo;
    const uint32_t cnt = counts[b];
    const bool single = cnt <= 2 * cap;
    const uint32_t first = offsets[b] + (t - task_off[b]) * cap;
    const uint32_t last = (single || first + cap > offsets[b] + cnt) ? offsets[b] + cnt : first + cap;
    g1_xyzz acc = g1_xyzz::inf();
    for (uint32_t j = first; j < last; j++) {
        const uint32_t e = sorted[j];
        g1_affine q = msm_ld_affine(table + (e & 0x7fffffffu));
        if (e & 0x80000000u) q.y = neg(q.y);
        xyzz_madd(acc, q);
    }
    if (single) buckets[b] = acc; else task_out[t] = acc;
}

// One warp per multi-task bucket: lanes stride over the bucket's partial sums, then a
// shared-memory tree over the 32 lanes.
__global__ void __launch_bounds__(128) msm_merge_kernel(const uint32_t* multi, const uint32_t* counts,
                                                       const uint32_t* task_off, uint32_t B, uint32_t cap,
                                                       const g1_xyzz* task_out, g1_xyzz* buckets) {
    extern __shared__ uint4 smem_raw[];
    g1_xyzz* sm = reinterpret_cast<g1_xyzz*>(smem_raw);
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t m = blockIdx.x * (blockDim.x >> 5) + wib;
    if (m >= multi[B]) return;  // whole warp exits together
    const uint32_t b = multi[m];
    const uint32_t nt = msm_ntasks(counts[b], cap), t0 = task_off[b];
    g1_xyzz acc = g1_xyzz::inf();
    for (uint32_t j = lane; j < nt; j += 32) xyzz_add(acc, task_out[t0 + j]);
    g1_xyzz* w = sm + wib * 32;
    w[lane] = acc;
    __syncwarp();
    for (unsigned s = 16; s > 0; s >>= 1) {
        if (lane < s) {
            g1_xyzz o = w[lane + s];
            xyzz_add(acc, o);
            w[lane] = acc;
        }
        __syncwarp();
    }
    if (lane == 0) buckets[b] = acc;
}

// ---- bucket reduction: sum_b (b + 1) * B[b] = sum_j 2^j * S_j,  S_j = sum of the buckets whose
// weight v = b + 1 has bit j set.  The c plane sums are independent tree reductions and the 2^j
// factors are applied in parallel, so the dependent chain is ~ (8 + 7 + log chunks + c + log c)
// point operations instead of a running sum over every bucket.
static constexpr unsigned RED_T = 128, RED_L = 8;  // threads per block, buckets per thread

// k-th weight (k >= 0) with bit j set
__device__ __forceinline__ uint32_t weight_with_bit(uint32_t k, unsigned j) {
    return ((k >> j) << (j + 1)) | (1u << j) | (k & ((1u << j) - 1));
}

// grid (chunks, c): block (x, j) sums RED_T * RED_L selected buckets of plane j
__global__ void __launch_bounds__(RED_T) msm_plane_sum_kernel(const g1_xyzz* buckets, uint32_t B, uint32_t chunks,
                                                            g1_xyzz* parts) {
    extern __shared__ uint4 smem_raw[];
    g1_xyzz* sm = reinterpret_cast<g1_xyzz*>(smem_raw);
    const unsigned tid = threadIdx.x, j = blockIdx.y;
    const uint32_t k0 = (blockIdx.x * RED_T + tid) * RED_L;
    g1_xyzz v = g1_xyzz::inf();
    for (unsigned i = 0; i < RED_L; i++) {
        const uint32_t wgt = weight_with_bit(k0 + i, j);
        if (wgt <= B) xyzz_add(v, buckets[wgt - 1]);
    }
    sm[tid] = v;
    __syncthreads();
    for (unsigned s = RED_T >> 1; s > 0; s >>= 1) {
        if (tid < s) {
            g1_xyzz o = sm[tid + s];
            xyzz_add(v, o);
            sm[tid] = v;
        }
        __syncthreads();
    }
    if (tid == 0) parts[(size_t)j * chunks + blockIdx.x] = v;
}

// block j (one warp): sums plane j's chunk partials, then lane 0 applies 2^j.
__global__ void __launch_bounds__(32) msm_plane_combine_kernel(const g1_xyzz* parts, uint32_t chunks,
                                                              g1_xyzz* planes) {
    __shared__ g1_xyzz w[32];
    const unsigned lane = threadIdx.x, j = blockIdx.x;
    g1_xyzz v = g1_xyzz::inf();
    for (uint32_t i = lane; i < chunks; i += 32) xyzz_add(v, parts[(size_t)j * chunks + i]);
    w[lane] = v;
    __syncwarp();
    for (unsigned s = 16; s > 0; s >>= 1) {
        if (lane < s) {
            g1_xyzz o = w[lane + s];
            xyzz_add(v, o);
            w[lane] = v;
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (unsigned i = 0; i < j; i++) xyzz_dbl(v);
        planes[j] = v;
    }
}

// one warp: tree over the c <= 32 weighted plane sums
__global__ void __launch_bounds__(32) msm_final_sum_kernel(const g1_xyzz* planes, unsigned c, g1_xyzz* out) {
    __shared__ g1_xyzz w[32];
    const unsigned lane = threadIdx.x;
    g1_xyzz v = lane < c ? planes[lane] : g1_xyzz::inf();
    w[lane] = v;
    __syncwarp();
    for (unsigned s = 16; s > 0; s >>= 1) {
        if (lane < s) {
            g1_xyzz o = w[lane + s];
            xyzz_add(v, o);
            w[lane] = v;
        }
        __syncwarp();
    }
    if (lane == 0) *out = v;
}

// T[w][i] = 2^c * T[w-1][i]: one thread per point walks the windows (load-time only).
__global__ void __launch_bounds__(128) srs_table_window_kernel(g1_affine* table, size_t n, unsigned c, unsigned W) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine p = msm_ld_affine(table + i);
    g1_xyzz acc = g1_xyzz::inf();
    xyzz_madd(acc, p);
    for (unsigned w = 1; w < W; w++) {
        for (unsigned j = 0; j < c; j++) xyzz_dbl(acc);
        table[(size_t)w * n + i] = xyzz_to_affine(acc);
    }
}

__global__ void msm_top_nonzero_kernel(const fr_t* scalars, size_t n, long long* top) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(scalars + i);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    if (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) atomicMax(top, (long long)i);
}

// ------------------------------------------------------------------ SRS generation
// tbl[w][d-1] = d * 2^(8w) * G, affine; 32 windows x 255 entries.
static constexpr unsigned FB_W = 32, FB_D = 255;

__device__ __forceinline__ g1_affine g1_generator() {
    // canonical coordinates of the standard BLS12-381 G1 generator
    const uint32_t gx[12] = {0xdb22c6bbu, 0xfb3af00au, 0xf97a1aefu, 0x6c55e83fu, 0x171bac58u, 0xa14e3a3fu,
                             0x9774b905u, 0xc3688c4fu, 0x4fa9ac0fu, 0x2695638cu, 0x3197d794u, 0x17f1d3a7u};
    const uint32_t gy[12] = {0x46c5e7e1u, 0x0caa2329u, 0xa2888ae4u, 0xd03cc744u, 0x2c04b3edu, 0x00db18cbu,
                             0xd5d00af6u, 0xfcf5e095u, 0x741d8ae4u, 0xa09e30edu, 0xe3aaa0f1u, 0x08b3f481u};
    g1_affine g;
    fq_t x, y;
#pragma unroll
    for (int i = 0; i < 12; i++) { x.l[i] = gx[i]; y.l[i] = gy[i]; }
    g.x = to_mont(x);
    g.y = to_mont(y);
    return g;
}

__global__ void srs_table_kernel(g1_affine* tbl) {
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= FB_W) return;
    g1_affine g = g1_generator();
    g1_xyzz base = g1_xyzz::inf();
    xyzz_madd(base, g);
    for (unsigned i = 0; i < 8 * w; i++) xyzz_dbl(base);
    g1_xyzz acc = g1_xyzz::inf();
    for (unsigned d = 0; d < FB_D; d++) {
        xyzz_add(acc, base);
        tbl[w * FB_D + d] = xyzz_to_affine(acc);
    }
}

__global__ void __launch_bounds__(128) srs_powers_kernel(const g1_affine* tbl, fr_t tau, size_t first, size_t n,
                                                        g1_affine* out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr_t s = from_mont(pow_u64(tau, (uint64_t)(first + i)));
    g1_xyzz acc = g1_xyzz::inf();
    for (unsigned w = 0; w < FB_W; w++) {
        const uint32_t d = (s.l[w >> 2] >> (8 * (w & 3))) & 255u;
        if (d) xyzz_madd(acc, msm_ld_affine(tbl + w * FB_D + d - 1));
    }
    out[i] = xyzz_to_affine(acc);
}

int srs_generate(zkp_ctx* ctx, const fr_t& tau, size_t first, size_t n, g1_affine* out_dev) {
    g1_affine* tbl = nullptr;
    ZKP_CUDA(ctx, cudaMalloc(&tbl, sizeof(g1_affine) * FB_W * FB_D));
    srs_table_kernel<<<1, 32, 0, ctx->stream>>>(tbl);
    ZKP_LAUNCHED(ctx);
    if (n) {
        srs_powers_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(tbl, tau, first, n, out_dev);
        ZKP_LAUNCHED(ctx);
    }
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZKP_CUDA(ctx, cudaFree(tbl));
    return ZKP_OK;
}

// ------------------------------------------------------------------ host driver
// Host-side Fq (6 x u64 CIOS Montgomery) for the one inversion per MSM.
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
static inline fq inv(const fq& a) {  // a^(p-2)
    uint64_t e[6]; memcpy(e, P, 48); e[0] -= 2;
    fq acc; memcpy(acc.l, ONE, 48);
    for (int i = 383; i >= 0; i--) {
        acc = mul(acc, acc);
        if ((e[i / 64] >> (i % 64)) & 1) acc = mul(acc, a);
    }
    return acc;
}
}  // namespace hostfq

// x = X / ZZ, y = Y / ZZZ with 1/ZZ = (ZZ / ZZZ)^2
static g1_affine host_xyzz_to_affine(const g1_xyzz& a) {
    g1_affine r = g1_affine::inf();
    if (a.is_inf()) return r;
    hostfq::fq X, Y, ZZ, ZZZ;
    memcpy(X.l, a.x.l, 48); memcpy(Y.l, a.y.l, 48); memcpy(ZZ.l, a.zz.l, 48); memcpy(ZZZ.l, a.zzz.l, 48);
    const hostfq::fq t = hostfq::inv(ZZZ);
    const hostfq::fq zi = hostfq::mul(ZZ, t);
    const hostfq::fq x = hostfq::mul(X, hostfq::mul(zi, zi)), y = hostfq::mul(Y, t);
    memcpy(r.x.l, x.l, 48); memcpy(r.y.l, y.l, 48);
    return r;
}

// Window width for an SRS of n powers: minimise (n * W mixed additions) + (bucket reduction,
// weighted for its lower parallel efficiency).
unsigned msm_choose_window(size_t n) {
    unsigned best = 4;
    double best_cost = 1e300;
    for (unsigned c = 4; c <= 16; c++) {
        const unsigned W = 255 / c + 1;
        const double cost = (double)(n ? n : 1) * W + 6.0 * (double)(1u << (c - 1));
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

// Fills rows 1 .. W-1 of the table from row 0 (the SRS powers themselves).
int srs_build_table(zkp_ctx* ctx, zkp_srs* srs) {
    if (srs->n == 0 || srs->W <= 1) return ZKP_OK;
    srs_table_window_kernel<<<(unsigned)((srs->n + 127) / 128), 128, 0, ctx->stream>>>(srs->d, srs->n, srs->c, srs->W);
    ZKP_LAUNCHED(ctx);
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

template <class T>
static int ensure(zkp_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (*cap >= need) return ZKP_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    ZKP_CUDA(ctx, cudaMalloc(p, need * sizeof(T)));
    *cap = need;
    return ZKP_OK;
}

static int msm_scratch(zkp_ctx* ctx, MsmScratch** out) {
    if (!ctx->msm) {
        ctx->msm = new MsmScratch();
        ZKP_CUDA(ctx, cudaMalloc(&ctx->msm->result, sizeof(g1_affine)));
        ZKP_CUDA(ctx, cudaMalloc(&ctx->msm->top, sizeof(long long)));
    }
    *out = ctx->msm;
    return ZKP_OK;
}

void msm_free(zkp_ctx* ctx) {
    MsmScratch* s = ctx->msm;
    if (!s) return;
    cudaFree(s->digits); cudaFree(s->sorted); cudaFree(s->counts); cudaFree(s->offsets); cudaFree(s->cursor);
    cudaFree(s->task_off); cudaFree(s->multi); cudaFree(s->buckets); cudaFree(s->task_out);
    cudaFree(s->part_a); cudaFree(s->part_b); cudaFree(s->result); cudaFree(s->top);
    delete s;
    ctx->msm = nullptr;
}

int msm_highest_nonzero(zkp_ctx* ctx, const fr_t* scalars_dev, size_t n, long long* out) {
    int rc;
    if ((rc = set_device(ctx))) return rc;
    MsmScratch* s;
    if ((rc = msm_scratch(ctx, &s))) return rc;
    long long* h = reinterpret_cast<long long*>(ctx->pinned);
    *h = -1;
    ZKP_CUDA(ctx, cudaMemcpyAsync(s->top, h, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
        msm_top_nonzero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(scalars_dev, n, s->top);
        ZKP_LAUNCHED(ctx);
    }
    ZKP_CUDA(ctx, cudaMemcpyAsync(h, s->top, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = *h;
    return ZKP_OK;
}

int msm_run(zkp_ctx* ctx, const zkp_srs* srs, const fr_t* scalars_dev, size_t n, g1_affine* out_host) {
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (n == 0) {
        memset(out_host, 0, sizeof(g1_affine));
        return ZKP_OK;
    }
    const unsigned c = srs->c, W = srs->W;
    if (n > srs->n || (size_t)W * srs->n >= (1ull << 31)) return ZKP_ERR_INVALID;
    ctx->msm_points += n;
    MsmScratch* s;
    if ((rc = msm_scratch(ctx, &s))) return rc;

    const uint32_t B = 1u << (c - 1);
    const size_t E = (size_t)W * n;  // upper bound on the number of non-zero digits
    // task size: >= 32 entries, and few enough tasks that the partial-sum array stays small
    const uint32_t cap = (uint32_t)((E >> 19) > 32 ? (E >> 19) : 32);
    const size_t max_tasks = E / cap + B + 1;
    // plane sums: weights with bit j set number at most B/2 (+1 for the top plane)
    const uint32_t chunks = (uint32_t)((B / 2 + 1 + RED_T * RED_L - 1) / (RED_T * RED_L));
    const size_t nparts = (size_t)chunks * c;

    if ((rc = ensure(ctx, &s->digits, &s->cap_entries, E))) return rc;
    if ((rc = ensure(ctx, &s->sorted, &s->cap_sorted, E))) return rc;
    if (s->cap_buckets < B) {
        size_t c1 = s->cap_buckets, c2 = c1, c3 = c1, c6 = c1;
        size_t c4 = c1 ? c1 + 1 : 0, c5 = c4;
        if ((rc = ensure(ctx, &s->counts, &c1, B))) return rc;
        if ((rc = ensure(ctx, &s->offsets, &c2, B))) return rc;
        if ((rc = ensure(ctx, &s->cursor, &c3, B))) return rc;
        if ((rc = ensure(ctx, &s->task_off, &c4, (size_t)B + 1))) return rc;
        if ((rc = ensure(ctx, &s->multi, &c5, (size_t)B + 1))) return rc;
        if ((rc = ensure(ctx, &s->buckets, &c6, B))) return rc;
        s->cap_buckets = B;
    }
    if ((rc = ensure(ctx, &s->task_out, &s->cap_tasks, max_tasks))) return rc;
    if (s->cap_partials < nparts) {
        size_t c1 = s->cap_partials, c2 = s->part_b ? 1 : 0;
        if ((rc = ensure(ctx, &s->part_a, &c1, nparts))) return rc;
        if ((rc = ensure(ctx, &s->part_b, &c2, 33))) return rc;
        s->cap_partials = nparts;
    }

    cudaStream_t st = ctx->stream;
    {
    ProfScope prof(ctx, "msm_sort");
    ZKP_CUDA(ctx, cudaMemsetAsync(s->counts, 0, B * sizeof(uint32_t), st));
    ZKP_CUDA(ctx, cudaMemsetAsync(s->buckets, 0, B * sizeof(g1_xyzz), st));  // all-zero XYZZ = infinity
    msm_digits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scalars_dev, n, c, W, s->digits, s->counts);
    ZKP_LAUNCHED(ctx);
    msm_scan_kernel<<<1, 1024, 0, st>>>(s->counts, s->offsets, s->cursor, s->task_off, s->multi, B, cap);
    ZKP_LAUNCHED(ctx);
    msm_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s->digits, n, W, (uint32_t)srs->n, s->cursor,
                                                                    s->sorted);
    ZKP_LAUNCHED(ctx);
    }
    {
    ProfScope prof(ctx, "msm_accumulate");
    msm_accumulate_kernel<<<(unsigned)((max_tasks + 127) / 128), 128, 0, st>>>(
        srs->d, s->sorted, s->offsets, s->counts, s->task_off, B, cap, s->buckets, s->task_out);
    ZKP_LAUNCHED(ctx);
    }
    {
    ProfScope prof(ctx, "msm_reduce");
    {
        // at most min(B, E / (2 cap)) buckets can consist of more than one task
        const size_t mm = E / (2 * (size_t)cap) < B ? E / (2 * (size_t)cap) : B;
        if (mm) {
            msm_merge_kernel<<<(unsigned)((mm + 3) / 4), 128, 128 * sizeof(g1_xyzz), st>>>(
                s->multi, s->counts, s->task_off, B, cap, s->task_out, s->buckets);
            ZKP_LAUNCHED(ctx);
        }
    }
    msm_plane_sum_kernel<<<dim3(chunks, c), RED_T, RED_T * sizeof(g1_xyzz), st>>>(s->buckets, B, chunks, s->part_a);
    ZKP_LAUNCHED(ctx);
    msm_plane_combine_kernel<<<c, 32, 0, st>>>(s->part_a, chunks, s->part_b);
    ZKP_LAUNCHED(ctx);
    msm_final_sum_kernel<<<1, 32, 0, st>>>(s->part_b, c, s->part_b + 32);
    ZKP_LAUNCHED(ctx);
    }
    // the single inversion of the conversion to affine runs on the host (one Fq Fermat chain
    // would occupy one GPU thread for ~0.6 ms)
    g1_xyzz* h = reinterpret_cast<g1_xyzz*>(ctx->pinned);
    ZKP_CUDA(ctx, cudaMemcpyAsync(h, s->part_b + 32, sizeof(g1_xyzz), cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(ctx, cudaStreamSynchronize(st));
    *out_host = host_xyzz_to_affine(*h);
    return ZKP_OK;
}

}  // namespace zkp

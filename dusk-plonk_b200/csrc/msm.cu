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
//   2. msm_scan_*_kernel      bucket offsets (three-launch tiled exclusive scan)
//   3. msm_scatter_kernel     counting sort of (table index, sign) by bucket
//   3b. msm_affine_round_kernel (large jobs)  batched-affine pairwise rounds: every bucket's entries are
//                             added in pairs, level after level, in AFFINE coordinates -- 6 Fq
//                             multiplications per addition instead of 10, the one inversion per
//                             thread (binary GCD, ~88 multiplication times) shared by its K pairs
//   4. msm_accumulate_kernel  one thread per chunk of L consecutive sorted entries, whatever
//                             buckets they fall in: XYZZ mixed additions (8M + 2S each, 384-bit
//                             Montgomery on the integer pipe) -- the hot kernel; equal work per
//                             thread for ANY digit distribution (top window, tiny coefficients,
//                             carry-only digits)
//   5. msm_merge[_giant]_kernel  buckets cut by chunk boundaries: add their partial sums
//   6. msm_rowcol / weighted_planes / final_sum    sum_b (b+1) * B[b]: row + column sums of the
//                             (hi, lo) weight grid, then two short bit-plane weighted sums
//   7. host: one binary-GCD inversion per commit group (host_inv.h) -> affine
// Addition in G1 is commutative and the result is normalised to affine, so the output is
// bit-identical to any correct CPU evaluation regardless of accumulation order.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "comm.cuh"
#include "fq_inv32.cuh"
#include "host_fq.h"
#include "host_inv.h"

namespace zkp {

static constexpr unsigned MSM_MAX_BATCH = 8;   // polynomials committed by one set of launches
static constexpr unsigned GIANT_PARTS = 64;    // buckets with more partial sums get a whole block

struct MsmBatch {
    const fr_t* sc[MSM_MAX_BATCH];
    uint32_t len[MSM_MAX_BATCH];
    uint32_t off[MSM_MAX_BATCH];   // SRS index of the polynomial's first scalar (a rank's range of a sharded commit)
};

struct MsmScratch {
    size_t cap_entries = 0, cap_sorted = 0, cap_buckets = 0, cap_planes = 0, cap_slots = 0;
    uint32_t* digits = nullptr;   // [nb][W * n]  bucket | sign << 31, 0xffffffff = zero digit
    uint32_t* sorted = nullptr;   // [nb][W * n]  table index | sign << 31, grouped by bucket
    uint32_t* counts = nullptr;   // [nb][B]
    uint32_t* offsets = nullptr;  // [levels][nb][B]  level r: offsets of ceil(count / 2^r) (level 0 = the sort's)
    uint32_t* cursor = nullptr;   // [nb][B]
    uint32_t* giant = nullptr;    // [nb][B]      buckets whose merge needs a whole block
    uint32_t* meta = nullptr;     // [nb][4]      entries, giant count, overflow flag, pad
    uint32_t* tile_sums = nullptr;  // [levels][MSM_MAX_BATCH][1024] scan scratch
    uint32_t* totals = nullptr;     // [levels][MSM_MAX_BATCH] entries left at each level
    uint32_t* plan = nullptr;       // [nb][E/2 + B]  source position | single flag of every output of a round
    g1_affine* aff[2] = {nullptr, nullptr};  // ping-pong outputs of the batched-affine rounds
    size_t cap_plan = 0, cap_aff[2] = {0, 0}, cap_offsets = 0;
    g1_xyzz* buckets = nullptr;   // [nb][B]
    g1_xyzz* slots = nullptr;     // [nb][2 * chunks]  head / tail partial sums of each chunk
    g1_xyzz* planes = nullptr;    // [nb][c * plane chunks] + [nb][32] + [nb]
    long long* top = nullptr;
    int acc_blocks_per_sm = 0;
    int acc_variant = 3;
    bool acc_variant_forced = false;
    int acc_blocks_per_sm2 = 0;   // occupancy of the 2-blocks/SM build used for small jobs
    int aff_blocks_per_sm = 0;    // occupancy of the batched-affine round kernel
    int aff_rounds_forced = -1;   // ZKP_MSM_AFFINE_ROUNDS: tuning / A-B knob (0 = XYZZ only)
    size_t aff_kmax = 192;        // most pairs one thread (one inversion) takes per round (ZKP_MSM_AFFINE_KMAX)
    size_t aff_min_pairs = (size_t)1 << 22;   // a round must have this many outputs to beat XYZZ (ZKP_MSM_AFFINE_MIN)
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

// Table entry: one affine point in a 128-byte slot, x in the first 64-byte DRAM atom and y in the second.
// The gathers of the accumulation are random single-point reads; a 96-byte point straddles atoms (measured:
// ~200 bytes of DRAM traffic per point read), whereas a padded entry costs exactly two atoms -- and exactly
// one when only x is needed (the forward pass of the batched-affine rounds).
__device__ __forceinline__ g1_affine msm_ld_tab(const g1_tab* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    g1_affine r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const uint4 v = __ldg(q + i), u = __ldg(q + 4 + i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        w[12 + 4 * i] = u.x; w[12 + 4 * i + 1] = u.y; w[12 + 4 * i + 2] = u.z; w[12 + 4 * i + 3] = u.w;
    }
    return r;
}
__device__ __forceinline__ void msm_st_tab(g1_tab* p, const g1_affine& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        q[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        q[4 + i] = make_uint4(w[12 + 4 * i], w[12 + 4 * i + 1], w[12 + 4 * i + 2], w[12 + 4 * i + 3]);
    }
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

// digits[b][w * n + i]: bucket | sign for scalar i of polynomial b, window w.  Coefficients at
// indices >= srs_n must be zero (PlonkParams::commit's degree check): a non-zero one raises the
// polynomial's overflow flag instead of producing entries.
__global__ void msm_digits_kernel(const __grid_constant__ MsmBatch batch, uint32_t n, uint32_t srs_n, unsigned c,
                                  unsigned W, uint32_t* digits, uint32_t* counts, uint32_t* meta) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned pb = blockIdx.y;
    if (i >= batch.len[pb]) return;
    const fr_t sm = msm_ld_fr(batch.sc[pb] + i);
    if ((uint64_t)batch.off[pb] + i >= srs_n) {
        if (!sm.is_zero()) meta[4 * pb + 2] = 1u;
        return;
    }
    const fr_t s = from_mont(sm);
    const uint32_t B = 1u << (c - 1);
    digits += (size_t)pb * W * n;
    counts += (size_t)pb * B;
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

// Exclusive scan of the B bucket counts of every polynomial -> entry offsets (and the scatter
// cursors), meta[4 pb] = number of entries.  Three small launches: per-tile sums, scan of the tile
// sums, per-tile rescan -- all loads and stores coalesced 16-byte accesses.
static constexpr unsigned SCAN_T = 256, SCAN_PER = 8, SCAN_TILE = SCAN_T * SCAN_PER;  // 2048 counters / block

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* ws, uint32_t* total) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += x;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t e = lane < nw ? ws[lane] : 0, x = e;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        ws[lane] = x - e;
        if (lane == 31) ws[32] = x;
    }
    __syncthreads();
    *total = ws[32];
    return ws[wid] + incl - v;
}

// Levels: level r scans ceil(count / 2^r) -- the bucket sizes after r batched-affine pairing rounds
// (level 0 is the sort itself).  All levels of all polynomials share the three launches:
// blockIdx.y = level * nb + polynomial.
static constexpr unsigned MSM_MAX_LEVELS = 9;   // level 0 + up to 8 pairing rounds

__device__ __forceinline__ uint32_t level_count(uint32_t c, unsigned lvl) { return (c + ((1u << lvl) - 1u)) >> lvl; }

// grid (tiles, levels * nb): tile_sums[level][pb][tile] = sum of the tile's counters at that level
__global__ void __launch_bounds__(SCAN_T) msm_scan_tiles_kernel(const uint32_t* counts, uint32_t B, uint32_t tiles,
                                                               unsigned nb, uint32_t* tile_sums) {
    __shared__ uint32_t ws[33];
    const unsigned pb = blockIdx.y % nb, lvl = blockIdx.y / nb;
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER;
    const uint32_t* c = counts + (size_t)pb * B;
    uint32_t sum = 0;
    if (base + SCAN_PER <= B) {
        const uint4 a = *reinterpret_cast<const uint4*>(c + base), b2 = *reinterpret_cast<const uint4*>(c + base + 4);
        sum = level_count(a.x, lvl) + level_count(a.y, lvl) + level_count(a.z, lvl) + level_count(a.w, lvl) +
              level_count(b2.x, lvl) + level_count(b2.y, lvl) + level_count(b2.z, lvl) + level_count(b2.w, lvl);
    } else {
        for (uint32_t i = base; i < B && i < base + SCAN_PER; i++) sum += level_count(c[i], lvl);
    }
    uint32_t total;
    block_exclusive_scan(sum, ws, &total);
    if (threadIdx.x == 0) tile_sums[(size_t)blockIdx.y * tiles + blockIdx.x] = total;
}

// one block per (level, polynomial): exclusive scan of its <= 1024 tile sums in place;
// totals[level][pb] = entries at that level, meta[4 pb] = entries of the sort (level 0)
__global__ void __launch_bounds__(1024) msm_scan_sums_kernel(uint32_t* tile_sums, uint32_t tiles, unsigned nb,
                                                            uint32_t* totals, uint32_t* meta) {
    __shared__ uint32_t ws[33];
    const unsigned pb = blockIdx.x % nb, lvl = blockIdx.x / nb;
    uint32_t* t = tile_sums + (size_t)blockIdx.x * tiles;
    const uint32_t v = threadIdx.x < tiles ? t[threadIdx.x] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, ws, &total);
    if (threadIdx.x < tiles) t[threadIdx.x] = ex;
    if (threadIdx.x == 0) {
        totals[lvl * MSM_MAX_BATCH + pb] = total;
        if (lvl == 0) { meta[4 * pb] = total; meta[4 * pb + 1] = 0; }
    }
}

// grid (tiles, levels * nb): offsets[level][pb][.] of the tile's buckets (+ the scatter cursors at level 0)
__global__ void __launch_bounds__(SCAN_T) msm_scan_apply_kernel(const uint32_t* counts, uint32_t B, uint32_t tiles,
                                                               unsigned nb, const uint32_t* tile_sums,
                                                               uint32_t* offsets, uint32_t* cursor) {
    __shared__ uint32_t ws[33];
    const unsigned pb = blockIdx.y % nb, lvl = blockIdx.y / nb;
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER;
    const uint32_t* c = counts + (size_t)pb * B;
    uint32_t v[SCAN_PER];
    uint32_t sum = 0;
#pragma unroll
    for (unsigned i = 0; i < SCAN_PER; i++) {
        v[i] = base + i < B ? level_count(c[base + i], lvl) : 0;
        sum += v[i];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, ws, &total) + tile_sums[(size_t)blockIdx.y * tiles + blockIdx.x];
    uint32_t* o = offsets + (size_t)blockIdx.y * B;
    uint32_t* cu = cursor + (size_t)pb * B;
#pragma unroll
    for (unsigned i = 0; i < SCAN_PER; i++) {
        if (base + i < B) {
            o[base + i] = run;
            if (lvl == 0) cu[base + i] = run;
        }
        run += v[i];
    }
}

// sorted[pos] = (w * stride + i) | sign, grouped by bucket
__global__ void msm_scatter_kernel(const __grid_constant__ MsmBatch batch, const uint32_t* digits, uint32_t n,
                                   unsigned W, uint32_t stride, uint32_t B, uint32_t* cursor, uint32_t* sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned pb = blockIdx.y;
    // (stride is the SRS length: scalars that meet no SRS power produced no digits -- msm_digits_kernel)
    if (i >= n || i >= batch.len[pb] || (uint64_t)batch.off[pb] + i >= stride) return;
    digits += (size_t)pb * W * n; sorted += (size_t)pb * W * n; cursor += (size_t)pb * B;
    for (unsigned w = 0; w < W; w++) {
        const uint32_t enc = digits[(size_t)w * n + i];
        if (enc == DIGIT_ZERO) continue;
        const uint32_t pos = atomicAdd(&cursor[enc & 0x7fffffffu], 1u);
        sorted[pos] = (w * stride + batch.off[pb] + i) | (enc & 0x80000000u);
    }
}

// ---- batched-affine pairing rounds ------------------------------------------------------------------
// Round r halves every bucket: output j of bucket b is the sum of its entries 2j and 2j + 1 of level r - 1
// (a bucket with an odd count passes its last entry through).  Affine addition costs
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1
// i.e. 3 multiplications once 1 / (x2 - x1) is known, and Montgomery's trick turns the K inversions of a
// thread into 3 (K - 1) multiplications plus ONE inversion: 6 per addition against 10 for the XYZZ mixed
// addition, plus 88 / K for the binary-GCD inversion (fq_inv32.cuh).  Entries of a bucket are contiguous in
// the sorted order at every level, so a round is: plan (where are my two inputs) -> forward pass (running
// product of the denominators, parked in the x slot of the output) -> inversion -> backward pass (the
// additions).  Thread t of a block owns outputs base + i * 128 + t: every global access of a warp is to 32
// neighbouring slots.  Exceptional pairs are exact: an operand at infinity (x = y = 0) passes the other
// through, equal points are doubled through the same inversion (denominator 2 y, numerator 3 x^2), opposite
// points give infinity.

// plan[q] = position of the first input of output q at the previous level | (1 << 31 if it has no partner).
// A thread plans 8 consecutive outputs: one binary search for the first, then it walks the (non-empty) buckets.
static constexpr unsigned PLAN_PER = 8;
__global__ void __launch_bounds__(256) msm_affine_plan_kernel(const uint32_t* offsets, const uint32_t* counts,
                                                             const uint32_t* totals, uint32_t B, unsigned nb,
                                                             unsigned r, size_t plan_stride, uint32_t* plan) {
    const unsigned pb = blockIdx.y;
    const uint32_t total = totals[r * MSM_MAX_BATCH + pb];
    const uint64_t q64 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * PLAN_PER;
    if (q64 >= total) return;
    const uint32_t q0 = (uint32_t)q64;
    const uint32_t* off_out = offsets + ((size_t)r * nb + pb) * B;
    const uint32_t* off_in = offsets + ((size_t)(r - 1) * nb + pb) * B;
    const uint32_t* cnt = counts + (size_t)pb * B;
    uint32_t blo = 0, bhi = B;      // the last bucket whose offset is <= q is never an empty one
    while (bhi - blo > 1) {
        const uint32_t mid = (blo + bhi) >> 1;
        if (off_out[mid] <= q0) blo = mid; else bhi = mid;
    }
    uint32_t b = blo, obeg = off_out[b], oend = obeg + level_count(cnt[b], r), ibeg = off_in[b];
    uint32_t cin = level_count(cnt[b], r - 1);
    uint32_t v[PLAN_PER];
    const uint32_t qe = q0 + PLAN_PER < total ? q0 + PLAN_PER : total;
#pragma unroll
    for (unsigned i = 0; i < PLAN_PER; i++) {
        const uint32_t q = q0 + i;
        v[i] = 0;
        if (q >= qe) continue;
        if (q >= oend) {
            do { b++; } while (cnt[b] == 0);
            obeg = off_out[b]; oend = obeg + level_count(cnt[b], r); ibeg = off_in[b];
            cin = level_count(cnt[b], r - 1);
        }
        const uint32_t j = q - obeg;
        v[i] = (ibeg + 2 * j) | ((2 * j + 1 >= cin) ? 0x80000000u : 0u);
    }
    uint32_t* out = plan + (size_t)pb * plan_stride + q0;
    if (q0 + PLAN_PER <= total && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
        reinterpret_cast<uint4*>(out)[0] = make_uint4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<uint4*>(out)[1] = make_uint4(v[4], v[5], v[6], v[7]);
    } else {
        for (unsigned i = 0; i < PLAN_PER && q0 + i < total; i++) out[i] = v[i];
    }
}

__device__ __forceinline__ fq_t msm_ld_fq(const fq_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    fq_t r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const uint4 v = q[i];
        r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
    }
    return r;
}
__device__ __forceinline__ void msm_st_fq(fq_t* p, const fq_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = make_uint4(v.l[4 * i], v.l[4 * i + 1], v.l[4 * i + 2], v.l[4 * i + 3]);
}

// Montgomery-domain inverse through the binary GCD: inv_mod gives (a R)^-1; times R^3 (Montgomery) = R / a
__device__ __noinline__ fq_t fq_inverse_bingcd(const fq_t& a) {
    const uint32_t r3w[12] = {0xd94ca1e0u, 0xed48ac6bu, 0x03a7adf8u, 0x315f831eu, 0x615e29ddu, 0x9a53352au,
                              0x921e1761u, 0x34c04e5eu, 0x65724728u, 0x2512d435u, 0x91755d4du, 0x0aa63460u};
    fq_t m = fq_t::modulus(), x, r3;
#pragma unroll
    for (int i = 0; i < 12; i++) r3.l[i] = r3w[i];
    fqinv::inv_mod<12>(a.l, m.l, 0x7ffdu /* -p^-1 mod 2^15 */, 51, x.l);
    return x * r3;
}

// inputs of a round: FIRST -> table entries named by the sorted list (128-byte slots, y at chunk 4);
// later rounds -> the dense affine points of the previous level (96 bytes, y at chunk 3)
template <bool FIRST>
struct AffSource {
    const g1_tab* table; const uint32_t* sorted; const g1_affine* in;
    static constexpr int YOFF = FIRST ? 4 : 3;
};

enum { PAIR_ADD = 0, PAIR_DBL = 1, PAIR_TAKE1 = 2, PAIR_TAKE2 = 3, PAIR_INF = 4 };

// Operand staging.  The inputs of a pair sit behind a chain of dependent loads (plan -> sorted list -> table),
// and a thread alternates between that chain and ~3000 cycles of multiplications: left alone, a third of
// the issue slots are lost to the loads (ncu, first version of this kernel: long-scoreboard stalls 2.2 per
// issue, integer pipe 61 %).  Here the addresses run two pairs ahead in registers and the operands one pair
// ahead through cp.async into shared memory -- each thread stages only its own data ([stage][chunk][thread]
// of 16-byte chunks, conflict-free), so the pipeline needs no barrier and no extra registers:
//   chunks 0-5 first point (x, y), 6-11 second point, 12-14 running product before this pair.
static constexpr int AFF_CHUNKS = 15, AFF_STAGES = 2;
static constexpr size_t AFF_SMEM = (size_t)AFF_STAGES * AFF_CHUNKS * 128 * sizeof(uint4);

__device__ __forceinline__ void cp_async16(uint4* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

struct AffPair {                      // where the inputs of one output live (16-byte chunk pointers)
    const uint4 *a1, *a2;             // a2 == nullptr: a lone entry that passes through
    bool n1, n2;                      // negate y (sign of the digit; first round only)
};
// The address chain is itself pipelined so that no instruction consumes a load issued in the same
// iteration: plan word three pairs ahead, the two sorted-list entries it names two pairs ahead (first
// round only), pointers formed when the cp.async is issued.
struct AffIdx { uint32_t pl, e1, e2; };

template <bool FIRST>
__device__ __forceinline__ AffIdx aff_idx(const AffSource<FIRST>& src, uint32_t pl) {
    AffIdx r;
    r.pl = pl; r.e1 = r.e2 = 0;
    if (FIRST) {
        r.e1 = src.sorted[pl & 0x7fffffffu];
        if (!(pl & 0x80000000u)) r.e2 = src.sorted[pl + 1];
    }
    return r;
}
template <bool FIRST>
__device__ __forceinline__ AffPair aff_pair(const AffSource<FIRST>& src, const AffIdx& x) {
    AffPair p;
    const bool single = (x.pl & 0x80000000u) != 0;
    if (FIRST) {
        p.a1 = reinterpret_cast<const uint4*>(src.table + (x.e1 & 0x7fffffffu)); p.n1 = (x.e1 & 0x80000000u) != 0;
        p.a2 = single ? nullptr : reinterpret_cast<const uint4*>(src.table + (x.e2 & 0x7fffffffu));
        p.n2 = (x.e2 & 0x80000000u) != 0;
    } else {
        p.a1 = reinterpret_cast<const uint4*>(src.in + (x.pl & 0x7fffffffu)); p.n1 = false;
        p.a2 = single ? nullptr : p.a1 + 6; p.n2 = false;
    }
    return p;
}

__device__ __forceinline__ fq_t aff_ld(const uint4* st, int chunk) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const uint4 v = st[(chunk + i) * 128];
        r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
    }
    return r;
}

// kind of the pair and its denominator; the common case (two finite points, different x) reads no y
template <int YOFF>
__device__ __forceinline__ int aff_classify(const uint4* st, const AffPair& p, const fq_t& x1, const fq_t& x2,
                                            bool have_y, fq_t* d) {
    const bool z1 = x1.is_zero(), z2 = x2.is_zero(), eq = (x1 == x2);
    if (!(z1 || z2 || eq)) { *d = x2 - x1; return PAIR_ADD; }
    fq_t y1 = have_y ? aff_ld(st, 3) : msm_ld_fq(reinterpret_cast<const fq_t*>(p.a1 + YOFF));
    fq_t y2 = have_y ? aff_ld(st, 9) : msm_ld_fq(reinterpret_cast<const fq_t*>(p.a2 + YOFF));
    *d = fq_t::one();
    if (z1 && y1.is_zero()) return PAIR_TAKE2;
    if (z2 && y2.is_zero()) return PAIR_TAKE1;
    if (!eq) { *d = x2 - x1; return PAIR_ADD; }
    if (p.n1) y1 = neg(y1);
    if (p.n2) y2 = neg(y2);
    if (!(y1 == y2) || y1.is_zero()) return PAIR_INF;             // P + (-P), or a point of order two
    *d = dbl(y1);
    return PAIR_DBL;
}

// cp.async the x coordinates of a pair (forward pass) / everything the addition needs (backward pass)
__device__ __forceinline__ void aff_issue_x(uint4* st, const AffPair& p) {
    if (p.a2) {
#pragma unroll
        for (int c = 0; c < 3; c++) { cp_async16(st + c * 128, p.a1 + c); cp_async16(st + (6 + c) * 128, p.a2 + c); }
    }
    cp_async_commit();
}
template <int YOFF>
__device__ __forceinline__ void aff_issue_all(uint4* st, const AffPair& p, const fq_t* prev) {
#pragma unroll
    for (int c = 0; c < 3; c++) { cp_async16(st + c * 128, p.a1 + c); cp_async16(st + (3 + c) * 128, p.a1 + YOFF + c); }
    if (p.a2) {
#pragma unroll
        for (int c = 0; c < 3; c++) { cp_async16(st + (6 + c) * 128, p.a2 + c); cp_async16(st + (9 + c) * 128, p.a2 + YOFF + c); }
        if (prev) {
            const uint4* g3 = reinterpret_cast<const uint4*>(prev);
#pragma unroll
            for (int c = 0; c < 3; c++) cp_async16(st + (12 + c) * 128, g3 + c);
        }
    }
    cp_async_commit();
}

template <bool FIRST>
__global__ void __launch_bounds__(128, 3) msm_affine_round_kernel(const g1_tab* table, const uint32_t* sorted,
                                                                 size_t entry_stride, const g1_affine* in,
                                                                 size_t in_stride, g1_affine* out, size_t out_stride,
                                                                 const uint32_t* plan, size_t plan_stride,
                                                                 const uint32_t* totals_r, uint32_t K) {
    extern __shared__ uint4 aff_sm[];
    const unsigned pb = blockIdx.y, t = threadIdx.x;
    const uint32_t total = totals_r[pb];
    const uint64_t base64 = (uint64_t)blockIdx.x * K * 128u;
    if (base64 + t >= total) return;
    const uint32_t q0 = (uint32_t)base64 + t;
    uint32_t Kt = (total - q0 + 127u) / 128u;       // outputs q0 + 128 i below total
    if (Kt > K) Kt = K;
    const AffSource<FIRST> src = {table, sorted + (size_t)pb * entry_stride, in + (size_t)pb * in_stride};
    out += (size_t)pb * out_stride;
    plan += (size_t)pb * plan_stride;
    uint4* const sm = aff_sm + t;
    constexpr int STAGE = AFF_CHUNKS * 128;

    // ---- forward: running product of the denominators, parked in out[q].x
    fq_t acc = fq_t::one();
    {
        // pipeline registers: cur = pair i (operands landed), nxt = pair i + 1 (in flight), ix = entries of
        // pair i + 2, pl3 = plan word of pair i + 3
        AffPair cur = aff_pair<FIRST>(src, aff_idx<FIRST>(src, plan[q0])), nxt = cur;
        aff_issue_x(sm, cur);
        if (Kt > 1) nxt = aff_pair<FIRST>(src, aff_idx<FIRST>(src, plan[q0 + 128u]));
        AffIdx ix = {0x80000000u, 0, 0};
        if (Kt > 2) ix = aff_idx<FIRST>(src, plan[q0 + 256u]);
        uint32_t pl3 = Kt > 3 ? plan[q0 + 384u] : 0x80000000u;
        for (uint32_t i = 0; i < Kt; i++) {
            if (i + 1 < Kt) aff_issue_x(sm + ((i + 1) & 1u) * STAGE, nxt); else cp_async_commit();
            const AffPair nn = aff_pair<FIRST>(src, ix);             // pair i + 2 (entries loaded last iteration)
            if (i + 3 < Kt) ix = aff_idx<FIRST>(src, pl3);
            if (i + 4 < Kt) pl3 = plan[q0 + 128u * (i + 4)];
            cp_async_wait1();
            if (cur.a2) {
                const uint4* st = sm + (i & 1u) * STAGE;
                fq_t d;
                const int kind = aff_classify<AffSource<FIRST>::YOFF>(st, cur, aff_ld(st, 0), aff_ld(st, 6), false, &d);
                if (kind <= PAIR_DBL) acc = acc * d;
            }
            msm_st_fq(&out[q0 + 128u * i].x, acc);
            cur = nxt; nxt = nn;
        }
    }
    fq_t inv = fq_inverse_bingcd(acc);
    // ---- backward: 1 / d_i = inv * prefix_(i-1), inv <- inv * d_i, then the addition itself
    {
        AffPair cur = aff_pair<FIRST>(src, aff_idx<FIRST>(src, plan[q0 + 128u * (Kt - 1)])), nxt = cur;
        aff_issue_all<AffSource<FIRST>::YOFF>(sm + ((Kt - 1) & 1u) * STAGE, cur, Kt > 1 ? &out[q0 + 128u * (Kt - 2)].x : nullptr);
        if (Kt > 1) nxt = aff_pair<FIRST>(src, aff_idx<FIRST>(src, plan[q0 + 128u * (Kt - 2)]));
        AffIdx ix = {0x80000000u, 0, 0};
        if (Kt > 2) ix = aff_idx<FIRST>(src, plan[q0 + 128u * (Kt - 3)]);
        uint32_t pl3 = Kt > 3 ? plan[q0 + 128u * (Kt - 4)] : 0x80000000u;
        for (uint32_t i = Kt; i-- > 0;) {
            if (i >= 1) aff_issue_all<AffSource<FIRST>::YOFF>(sm + ((i - 1) & 1u) * STAGE, nxt, i >= 2 ? &out[q0 + 128u * (i - 2)].x : nullptr);
            else cp_async_commit();
            const AffPair nn = aff_pair<FIRST>(src, ix);             // pair i - 2
            if (i >= 3) ix = aff_idx<FIRST>(src, pl3);
            if (i >= 4) pl3 = plan[q0 + 128u * (i - 4)];
            cp_async_wait1();
            const uint4* st = sm + (i & 1u) * STAGE;
            g1_affine res;
            res.x = aff_ld(st, 0);
            res.y = aff_ld(st, 3);
            if (cur.n1) res.y = neg(res.y);
            if (cur.a2) {
                const fq_t x2 = aff_ld(st, 6);
                fq_t d;
                const int kind = aff_classify<AffSource<FIRST>::YOFF>(st, cur, res.x, x2, true, &d);
                if (kind <= PAIR_DBL) {
                    const fq_t prev = i ? aff_ld(st, 12) : fq_t::one();
                    const fq_t dinv = inv * prev;
                    inv = inv * d;
                    fq_t num;
                    if (kind == PAIR_ADD) {
                        fq_t y2 = aff_ld(st, 9);
                        if (cur.n2) y2 = neg(y2);
                        num = y2 - res.y;
                    } else {
                        const fq_t xx = sqr(res.x);
                        num = dbl(xx) + xx;
                    }
                    const fq_t lam = num * dinv;
                    const fq_t x3 = sqr(lam) - res.x - x2;
                    res.y = lam * (res.x - x3) - res.y;
                    res.x = x3;
                } else if (kind == PAIR_TAKE2) {
                    res.x = x2;
                    res.y = aff_ld(st, 9);
                    if (cur.n2) res.y = neg(res.y);
                } else if (kind == PAIR_INF) {
                    res = g1_affine::inf();
                }
            }
            uint4* o = reinterpret_cast<uint4*>(out + q0 + 128u * i);
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&res);
#pragma unroll
            for (int k = 0; k < 6; k++) o[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
            cur = nxt; nxt = nn;
        }
    }
}

// First round: the inputs are random table entries.  Staging them through cp.async was measured SLOWER here
// (15 uncoalesced 16-byte copies per pair keep the load/store unit's queue full: ncu mio_throttle 2.0 per issue,
// 12.1 ms against 9.6 ms at 2^22), so this round loads straight into registers: the x coordinates of the next
// pair one iteration ahead in the forward pass, the address chain (plan word -> sorted-list entries) two
// ahead in both passes.
__device__ __forceinline__ fq_t msm_ldg_fq(const uint4* q) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const uint4 v = __ldg(q + i);
        r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
    }
    return r;
}

__global__ void __launch_bounds__(128, 3) msm_affine_first_kernel(const g1_tab* table, const uint32_t* sorted,
                                                                 size_t entry_stride, g1_affine* out, size_t out_stride,
                                                                 const uint32_t* plan, size_t plan_stride,
                                                                 const uint32_t* totals_r, uint32_t K) {
    const unsigned pb = blockIdx.y, t = threadIdx.x;
    const uint32_t total = totals_r[pb];
    const uint64_t base64 = (uint64_t)blockIdx.x * K * 128u;
    if (base64 + t >= total) return;
    const uint32_t q0 = (uint32_t)base64 + t;
    uint32_t Kt = (total - q0 + 127u) / 128u;
    if (Kt > K) Kt = K;
    const AffSource<true> src = {table, sorted + (size_t)pb * entry_stride, nullptr};
    out += (size_t)pb * out_stride;
    plan += (size_t)pb * plan_stride;
    constexpr int YOFF = AffSource<true>::YOFF;

    fq_t acc = fq_t::one();
    {
        AffPair cur = aff_pair<true>(src, aff_idx<true>(src, plan[q0]));
        AffIdx ix = {0x80000000u, 0, 0};
        if (Kt > 1) ix = aff_idx<true>(src, plan[q0 + 128u]);
        uint32_t pl2 = Kt > 2 ? plan[q0 + 256u] : 0x80000000u;
        fq_t x1 = fq_t::zero(), x2 = fq_t::zero();
        if (cur.a2) { x1 = msm_ldg_fq(cur.a1); x2 = msm_ldg_fq(cur.a2); }
        for (uint32_t i = 0; i < Kt; i++) {
            const AffPair nxt = aff_pair<true>(src, ix);             // pair i + 1 (entries loaded last iteration)
            if (i + 2 < Kt) ix = aff_idx<true>(src, pl2);
            if (i + 3 < Kt) pl2 = plan[q0 + 128u * (i + 3)];
            fq_t nx1 = fq_t::zero(), nx2 = fq_t::zero();
            if (i + 1 < Kt && nxt.a2) { nx1 = msm_ldg_fq(nxt.a1); nx2 = msm_ldg_fq(nxt.a2); }
            if (cur.a2) {
                fq_t d;
                const int kind = aff_classify<YOFF>(nullptr, cur, x1, x2, false, &d);
                if (kind <= PAIR_DBL) acc = acc * d;
            }
            msm_st_fq(&out[q0 + 128u * i].x, acc);
            cur = nxt; x1 = nx1; x2 = nx2;
        }
    }
    fq_t inv = fq_inverse_bingcd(acc);
    {
        AffPair cur = aff_pair<true>(src, aff_idx<true>(src, plan[q0 + 128u * (Kt - 1)]));
        AffIdx ix = {0x80000000u, 0, 0};
        if (Kt > 1) ix = aff_idx<true>(src, plan[q0 + 128u * (Kt - 2)]);
        uint32_t pl2 = Kt > 2 ? plan[q0 + 128u * (Kt - 3)] : 0x80000000u;
        for (uint32_t i = Kt; i-- > 0;) {
            const AffPair nxt = aff_pair<true>(src, ix);             // pair i - 1
            if (i >= 2) ix = aff_idx<true>(src, pl2);
            if (i >= 3) pl2 = plan[q0 + 128u * (i - 3)];
            g1_affine res = msm_ld_tab(reinterpret_cast<const g1_tab*>(cur.a1));
            if (cur.n1) res.y = neg(res.y);
            if (cur.a2) {
                const fq_t x2 = msm_ldg_fq(cur.a2);
                const bool z1 = res.x.is_zero(), z2 = x2.is_zero(), eq = (res.x == x2);
                int kind = PAIR_ADD;
                fq_t y2 = msm_ldg_fq(cur.a2 + YOFF);
                if (cur.n2) y2 = neg(y2);
                if (z1 || z2 || eq) {                                // same decisions as aff_classify
                    if (z1 && res.y.is_zero()) kind = PAIR_TAKE2;
                    else if (z2 && y2.is_zero()) kind = PAIR_TAKE1;
                    else if (!eq) kind = PAIR_ADD;
                    else if (!(res.y == y2) || res.y.is_zero()) kind = PAIR_INF;
                    else kind = PAIR_DBL;
                }
                if (kind <= PAIR_DBL) {
                    const fq_t prev = i ? msm_ld_fq(&out[q0 + 128u * (i - 1)].x) : fq_t::one();
                    const fq_t dinv = inv * prev;
                    fq_t num;
                    if (kind == PAIR_ADD) {
                        inv = inv * (x2 - res.x);
                        num = y2 - res.y;
                    } else {
                        inv = inv * dbl(res.y);
                        const fq_t xx = sqr(res.x);
                        num = dbl(xx) + xx;
                    }
                    const fq_t lam = num * dinv;
                    const fq_t x3 = sqr(lam) - res.x - x2;
                    res.y = lam * (res.x - x3) - res.y;
                    res.x = x3;
                } else if (kind == PAIR_TAKE2) {
                    res.x = x2; res.y = y2;
                } else if (kind == PAIR_INF) {
                    res = g1_affine::inf();
                }
            }
            uint4* o = reinterpret_cast<uint4*>(out + q0 + 128u * i);
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&res);
#pragma unroll
            for (int k = 0; k < 6; k++) o[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
            cur = nxt;
        }
    }
}

// The hot kernel of the XYZZ path.  Thread t owns the L consecutive sorted entries [t L, (t+1) L) whatever
// buckets they belong to, so every thread does the same number of mixed additions for ANY digit
// distribution.  A bucket that lies inside one chunk is written directly; a bucket cut by a chunk
// boundary leaves partial sums in the chunk's head slot (first segment of the chunk) or tail slot
// (last segment), which msm_merge_kernel adds up.
// GATHER: entries are (table index, sign) pairs of the sort; otherwise they are the affine points the
// batched-affine rounds left at level `lvl` (bucket b then holds ceil(count / 2^lvl) of them).
// MB = resident blocks per SM the register allocation is bounded for (2: 174 registers, 3: 168).  Small,
// one-wave jobs run best with 2, large ones with 3; 4 .. 6 (128 .. 80 registers, spills) were measured
// slower.
template <int MB, bool GATHER>
__global__ void __launch_bounds__(128, MB) msm_accumulate_kernel(const void* points, const uint32_t* sorted,
                                                            const uint32_t* offsets, const uint32_t* counts,
                                                            const uint32_t* totals, unsigned lvl, uint32_t B, uint32_t L,
                                                            uint32_t nchunks, size_t entry_stride, g1_xyzz* buckets,
                                                            g1_xyzz* slots) {
    const unsigned pb = blockIdx.y;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = totals[pb];
    const uint64_t start64 = (uint64_t)t * L;
    if (start64 >= total) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = start + L < total ? start + L : total;
    const g1_tab* table = static_cast<const g1_tab*>(points);                       // GATHER: the window table
    const g1_affine* dense = static_cast<const g1_affine*>(points) + (GATHER ? 0 : (size_t)pb * entry_stride);
    if (GATHER) sorted += (size_t)pb * entry_stride;
    offsets += (size_t)pb * B; counts += (size_t)pb * B;
    buckets += (size_t)pb * B; slots += (size_t)pb * 2 * nchunks;
    // bucket holding entry `start`: the last b with offsets[b] <= start
    uint32_t blo = 0, bhi = B;
    while (bhi - blo > 1) {
        const uint32_t mid = (blo + bhi) >> 1;
        if (offsets[mid] <= start) blo = mid; else bhi = mid;
    }
    uint32_t bk = blo, bbeg = offsets[bk], bend = bbeg + level_count(counts[bk], lvl);
    uint32_t seg = start;
    bool first = true;
    g1_xyzz acc = g1_xyzz::inf();
    for (uint32_t j = start; j < end; j++) {
        if (j >= bend) {  // the bucket ended inside this chunk
            if (seg == bbeg) buckets[bk] = acc;          // ... and began inside it too: complete
            else slots[2 * t] = acc;                     // continuation from the previous chunk
            first = false;
            do { bk++; } while (counts[bk] == 0);
            bbeg = offsets[bk]; bend = bbeg + level_count(counts[bk], lvl);
            seg = j;
            acc = g1_xyzz::inf();
        }
        g1_affine q;
        if (GATHER) {
            const uint32_t e = sorted[j];
            q = msm_ld_tab(table + (e & 0x7fffffffu));
            if (e & 0x80000000u) q.y = neg(q.y);
        } else {
            q = msm_ld_affine(dense + j);
        }
        xyzz_madd(acc, q);
    }
    if (seg == bbeg && end == bend) buckets[bk] = acc;
    else slots[2 * t + (first ? 0 : 1)] = acc;
}

// One thread per bucket: empty -> infinity; inside one chunk -> already written; otherwise add the
// partial sums of the chunks it spans (buckets spanning > GIANT_PARTS chunks go to the block kernel).
__global__ void __launch_bounds__(128, 4) msm_merge_kernel(const uint32_t* offsets, const uint32_t* counts, unsigned lvl,
                                                       uint32_t B, uint32_t L, uint32_t nchunks, const g1_xyzz* slots,
                                                       g1_xyzz* buckets, uint32_t* giant, uint32_t* meta) {
    const unsigned pb = blockIdx.y;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t cnt = level_count(counts[(size_t)pb * B + b], lvl), beg = offsets[(size_t)pb * B + b];
    g1_xyzz* out = buckets + (size_t)pb * B + b;
    if (cnt == 0) { *out = g1_xyzz::inf(); return; }
    const uint32_t t0 = beg / L, t1 = (beg + cnt - 1) / L;
    if (t0 == t1) return;
    if (t1 - t0 + 1 > GIANT_PARTS) {
        giant[(size_t)pb * B + atomicAdd(&meta[4 * pb + 1], 1u)] = b;
        return;
    }
    slots += (size_t)pb * 2 * nchunks;
    g1_xyzz acc = slots[2 * t0 + (beg == t0 * L ? 0 : 1)];
    for (uint32_t t = t0 + 1; t <= t1; t++) xyzz_add(acc, slots[2 * t]);
    *out = acc;
}

// One block per giant bucket (e.g. a top window holding only the carry digit).
__global__ void __launch_bounds__(128) msm_merge_giant_kernel(const uint32_t* offsets, const uint32_t* counts,
                                                             unsigned lvl, uint32_t B, uint32_t L, uint32_t nchunks,
                                                             const g1_xyzz* slots, g1_xyzz* buckets,
                                                             const uint32_t* giant, const uint32_t* meta) {
    extern __shared__ uint4 smem_raw[];
    g1_xyzz* sm = reinterpret_cast<g1_xyzz*>(smem_raw);
    const unsigned pb = blockIdx.y, tid = threadIdx.x;
    if (blockIdx.x >= meta[4 * pb + 1]) return;
    const uint32_t b = giant[(size_t)pb * B + blockIdx.x];
    const uint32_t cnt = level_count(counts[(size_t)pb * B + b], lvl), beg = offsets[(size_t)pb * B + b];
    const uint32_t t0 = beg / L, t1 = (beg + cnt - 1) / L;
    slots += (size_t)pb * 2 * nchunks;
    g1_xyzz v = g1_xyzz::inf();
    for (uint32_t t = t0 + tid; t <= t1; t += blockDim.x)
        xyzz_add(v, slots[2 * t + ((t == t0 && beg != t0 * L) ? 1 : 0)]);
    sm[tid] = v;
    __syncthreads();
    for (unsigned s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (tid < s) {
            g1_xyzz o = sm[tid + s];
            xyzz_add(v, o);
            sm[tid] = v;
        }
        __syncthreads();
    }
    if (tid == 0) buckets[(size_t)pb * B + b] = v;
}

// ---- bucket reduction: sum_v v * Bk[v], v = b + 1 in [1, B].  Split v = hi * 2^h + lo:
//     sum_v v Bk[v] = 2^h * sum_hi hi * R[hi] + sum_lo lo * C[lo]
// with R[hi] / C[lo] the row / column sums of the (hi, lo) grid -- 2 B additions in total, all
// rows and columns in parallel -- and the two short weighted sums done by bit planes
// (sum_x x A[x] = sum_j 2^j * sum_{x: bit j} A[x]), the 2^j factors applied in parallel.
// ---- cooperative point arithmetic ----------------------------------------------------------------
// The reduction is a dependency chain (~30 point additions + c doublings deep) on a mostly idle GPU,
// and one point addition is 14 dependent Fq multiplications when a single thread runs it.  Here a
// block of four warps -- one per SM sub-partition -- adds 32 pairs of points at a time: warp w takes
// the w-th multiplication of each dependency level (add: 4 levels instead of 14 multiplications,
// double: 3 instead of 9), operands and intermediates live in shared memory in limb-major order
// (conflict-free), one barrier per level.  Exceptional lanes (an operand at infinity, equal or
// opposite points) are flagged and redone by warp 0 with the generic formulas, so results are exact.
enum { AX, AY, AZZ, AZZZ, BX, BY, BZZ, BZZZ, TU1, TU2, TS1, TS2, TP, TPP, TR, TRR, TZZ, TZZZ, TPPP, TQ, TT, TV, NSLOT };
struct CoopSm {
    uint32_t s[NSLOT][12][32];
    uint32_t flag[32];
};

__device__ __forceinline__ fq_t cs_ld(const CoopSm& sm, int slot, unsigned l) {
    fq_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = sm.s[slot][i][l];
    return r;
}
__device__ __forceinline__ void cs_st(CoopSm& sm, int slot, unsigned l, const fq_t& v) {
#pragma unroll
    for (int i = 0; i < 12; i++) sm.s[slot][i][l] = v.l[i];
}
__device__ __forceinline__ g1_xyzz cs_ld_point(const CoopSm& sm, int base, unsigned l) {
    g1_xyzz r;
    r.x = cs_ld(sm, base, l); r.y = cs_ld(sm, base + 1, l); r.zz = cs_ld(sm, base + 2, l); r.zzz = cs_ld(sm, base + 3, l);
    return r;
}
__device__ __forceinline__ void cs_st_point(CoopSm& sm, int base, unsigned l, const g1_xyzz& v) {
    cs_st(sm, base, l, v.x); cs_st(sm, base + 1, l, v.y); cs_st(sm, base + 2, l, v.zz); cs_st(sm, base + 3, l, v.zzz);
}
// warp w fetches coordinate w of lane l's point (nullptr = infinity) into slot base + w
__device__ __forceinline__ void cs_fetch(CoopSm& sm, int base, unsigned w, unsigned l, const g1_xyzz* p) {
    if (p) {
        const uint4* q = reinterpret_cast<const uint4*>(p) + 3 * w;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const uint4 v = q[i];
            sm.s[base + w][4 * i][l] = v.x; sm.s[base + w][4 * i + 1][l] = v.y;
            sm.s[base + w][4 * i + 2][l] = v.z; sm.s[base + w][4 * i + 3][l] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 12; i++) sm.s[base + w][i][l] = 0;
    }
}
__device__ __forceinline__ void cs_emit(const CoopSm& sm, int base, unsigned w, unsigned l, g1_xyzz* p) {
    uint4* q = reinterpret_cast<uint4*>(p) + 3 * w;
#pragma unroll
    for (int i = 0; i < 3; i++)
        q[i] = make_uint4(sm.s[base + w][4 * i][l], sm.s[base + w][4 * i + 1][l], sm.s[base + w][4 * i + 2][l],
                          sm.s[base + w][4 * i + 3][l]);
}

// exceptional lanes only (rare): kept out of line so the common path stays within its register budget
__device__ __noinline__ void coop_add_slow(CoopSm& sm, unsigned l) {
    g1_xyzz a = cs_ld_point(sm, AX, l);
    const g1_xyzz b = cs_ld_point(sm, BX, l);
    xyzz_add(a, b);
    cs_st_point(sm, AX, l, a);
}
__device__ __noinline__ void coop_dbl_slow(CoopSm& sm, unsigned l) {
    g1_xyzz a = cs_ld_point(sm, AX, l);
    xyzz_dbl(a);
    cs_st_point(sm, AX, l, a);
}

// A[l] <- A[l] + B[l] for the lanes whose threads pass on = true (the four warps agree per lane).
// Called by all 128 threads; ends with a barrier.
__device__ __noinline__ void coop_add(CoopSm& sm, bool on) {
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (on) {
        if (w == 0) {
            const fq_t azz = cs_ld(sm, AZZ, l), bzz = cs_ld(sm, BZZ, l);
            sm.flag[l] = (azz.is_zero() ? 1u : 0u) | (bzz.is_zero() ? 2u : 0u);
            cs_st(sm, TU1, l, cs_ld(sm, AX, l) * bzz);
        } else if (w == 1) {
            cs_st(sm, TU2, l, cs_ld(sm, BX, l) * cs_ld(sm, AZZ, l));
        } else if (w == 2) {
            cs_st(sm, TS1, l, cs_ld(sm, AY, l) * cs_ld(sm, BZZZ, l));
        } else {
            cs_st(sm, TS2, l, cs_ld(sm, BY, l) * cs_ld(sm, AZZZ, l));
        }
    }
    __syncthreads();
    if (on) {
        if (w == 0) {
            const fq_t P = cs_ld(sm, TU2, l) - cs_ld(sm, TU1, l);
            if (P.is_zero()) sm.flag[l] |= 4u;
            cs_st(sm, TP, l, P);
            cs_st(sm, TPP, l, sqr(P));
        } else if (w == 1) {
            const fq_t R = cs_ld(sm, TS2, l) - cs_ld(sm, TS1, l);
            cs_st(sm, TR, l, R);
            cs_st(sm, TRR, l, sqr(R));
        } else if (w == 2) {
            cs_st(sm, TZZ, l, cs_ld(sm, AZZ, l) * cs_ld(sm, BZZ, l));
        } else {
            cs_st(sm, TZZZ, l, cs_ld(sm, AZZZ, l) * cs_ld(sm, BZZZ, l));
        }
    }
    __syncthreads();
    const bool fast = on && sm.flag[l] == 0;
    if (fast) {
        if (w == 0) cs_st(sm, TPPP, l, cs_ld(sm, TP, l) * cs_ld(sm, TPP, l));
        else if (w == 1) cs_st(sm, TQ, l, cs_ld(sm, TU1, l) * cs_ld(sm, TPP, l));
        else if (w == 2) cs_st(sm, AZZ, l, cs_ld(sm, TZZ, l) * cs_ld(sm, TPP, l));
    }
    __syncthreads();
    if (fast) {
        if (w == 0) {
            const fq_t Q = cs_ld(sm, TQ, l);
            const fq_t X3 = cs_ld(sm, TRR, l) - cs_ld(sm, TPPP, l) - dbl(Q);
            cs_st(sm, TT, l, cs_ld(sm, TR, l) * (Q - X3));
            cs_st(sm, AX, l, X3);
        } else if (w == 1) {
            cs_st(sm, TV, l, cs_ld(sm, TS1, l) * cs_ld(sm, TPPP, l));
        } else if (w == 2) {
            cs_st(sm, AZZZ, l, cs_ld(sm, TZZZ, l) * cs_ld(sm, TPPP, l));
        }
    }
    __syncthreads();
    if (w == 0 && on) {
        if (fast) {
            cs_st(sm, AY, l, cs_ld(sm, TT, l) - cs_ld(sm, TV, l));
        } else {  // infinity / doubling / cancellation: generic formulas on the untouched operands
            coop_add_slow(sm, l);
        }
    }
    __syncthreads();
}

// A[l] <- 2 A[l] for the lanes with on = true.
__device__ __noinline__ void coop_dbl(CoopSm& sm, bool on) {
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (on) {
        if (w == 0) {
            const fq_t y = cs_ld(sm, AY, l);
            sm.flag[l] = (cs_ld(sm, AZZ, l).is_zero() ? 1u : 0u) | (y.is_zero() ? 2u : 0u);
            const fq_t U = dbl(y);
            cs_st(sm, TU1, l, U);
            cs_st(sm, TPP, l, sqr(U));             // V
        } else if (w == 1) {
            const fq_t X2 = sqr(cs_ld(sm, AX, l));
            cs_st(sm, TR, l, dbl(X2) + X2);        // M
        }
    }
    __syncthreads();
    const bool fast = on && sm.flag[l] == 0;
    if (fast) {
        if (w == 0) cs_st(sm, TPPP, l, cs_ld(sm, TU1, l) * cs_ld(sm, TPP, l));        // W
        else if (w == 1) cs_st(sm, TQ, l, cs_ld(sm, AX, l) * cs_ld(sm, TPP, l));      // S
        else if (w == 2) cs_st(sm, TRR, l, sqr(cs_ld(sm, TR, l)));                    // M^2
        else cs_st(sm, AZZ, l, cs_ld(sm, TPP, l) * cs_ld(sm, AZZ, l));
    }
    __syncthreads();
    if (fast) {
        if (w == 0) {
            const fq_t S = cs_ld(sm, TQ, l);
            const fq_t X3 = cs_ld(sm, TRR, l) - dbl(S);
            cs_st(sm, TT, l, cs_ld(sm, TR, l) * (S - X3));
            cs_st(sm, AX, l, X3);
        } else if (w == 1) {
            cs_st(sm, TV, l, cs_ld(sm, TPPP, l) * cs_ld(sm, AY, l));
        } else if (w == 2) {
            cs_st(sm, AZZZ, l, cs_ld(sm, TPPP, l) * cs_ld(sm, AZZZ, l));
        }
    }
    __syncthreads();
    if (w == 0 && on) {
        if (fast) {
            cs_st(sm, AY, l, cs_ld(sm, TT, l) - cs_ld(sm, TV, l));
        } else {
            coop_dbl_slow(sm, l);
        }
    }
    __syncthreads();
}

// The block's 32 lanes are split into groups of `lpo` (a power of two): group o sums item(o, 0 .. len)
// into A[o * lpo].  Lane p of a group adds up items p, p + lpo, .. (every lane busy, no idle tree
// levels), then a log2(lpo)-level tree folds the group.  Fewer lanes per output = less idle tree work
// but a longer dependency chain: the launcher picks lpo from how full the GPU is.
// item(o, i) returns the point's address or nullptr for infinity; valid(o) = group o has an output.
template <class F, class V>
__device__ __forceinline__ void coop_group_sum(CoopSm& sm, unsigned lpo, uint32_t len, V valid, F item) {
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const unsigned o = l / lpo, p = l & (lpo - 1);
    const bool ok = valid(o);
    cs_fetch(sm, AX, w, l, (ok && p < len) ? item(o, p) : nullptr);
    __syncthreads();
    for (uint32_t i0 = lpo; i0 < len; i0 += lpo) {
        cs_fetch(sm, BX, w, l, (ok && i0 + p < len) ? item(o, i0 + p) : nullptr);
        __syncthreads();
        coop_add(sm, ok);
    }
    unsigned s = lpo >> 1;
    while (s >= len && s > 0) s >>= 1;   // lanes >= len hold infinity: skip the empty levels
    for (; s > 0; s >>= 1) {
        if (p < s) {
#pragma unroll
            for (int i = 0; i < 12; i++) sm.s[BX + w][i][l] = sm.s[AX + w][i][l + s];
        }
        __syncthreads();
        coop_add(sm, ok && p < s);
    }
}

// grid (ceil(nrows / R) + ceil(ncols / R), nb), 128 threads, R = 32 / lpo outputs per block:
// rc[pb][x] = row sum (x < nrows) or column sum of the weight grid
__global__ void __launch_bounds__(128, 4) msm_rowcol_coop_kernel(const g1_xyzz* buckets, uint32_t B, unsigned h,
                                                                 uint32_t nrows, uint32_t ncols, unsigned lpo,
                                                                 g1_xyzz* rc) {
    __shared__ CoopSm sm;
    const unsigned pb = blockIdx.y, R = 32 / lpo;
    const uint32_t row_blocks = (nrows + R - 1) / R;
    const bool is_row = blockIdx.x < row_blocks;
    const uint32_t x0 = is_row ? blockIdx.x * R : (blockIdx.x - row_blocks) * R;   // first row / column of the block
    const uint32_t count = is_row ? nrows : ncols, len = is_row ? ncols : nrows;
    buckets += (size_t)pb * B;
    coop_group_sum(sm, lpo, len,
        [&](unsigned o) { return x0 + o < count; },
        [&](unsigned o, uint32_t i) -> const g1_xyzz* {
            const uint32_t wgt = is_row ? ((x0 + o) << h) + i : (i << h) + (x0 + o);
            return (wgt >= 1 && wgt <= B) ? buckets + (wgt - 1) : nullptr;
        });
    const unsigned l = threadIdx.x & 31, o = l / lpo;
    if ((l & (lpo - 1)) == 0 && x0 + o < count)
        cs_emit(sm, AX, threadIdx.x >> 5, l, rc + (size_t)pb * (nrows + ncols) + (is_row ? 0 : nrows) + x0 + o);
}

// grid (planes_r + planes_c, nb), 128 threads: block j sums the rows (columns) whose index has bit j set
// and applies the plane's factor 2^(h + j) (2^j) by doublings.
__global__ void __launch_bounds__(128, 4) msm_planes_coop_kernel(const g1_xyzz* rc, unsigned h, uint32_t nrows,
                                                              uint32_t ncols, unsigned planes_r, g1_xyzz* planes) {
    __shared__ CoopSm sm;
    const unsigned pb = blockIdx.y;
    rc += (size_t)pb * (nrows + ncols);
    const bool is_row = blockIdx.x < planes_r;
    const unsigned j = is_row ? blockIdx.x : blockIdx.x - planes_r;
    const g1_xyzz* arr = is_row ? rc : rc + nrows;
    const uint32_t len = is_row ? nrows : ncols;
    // indices below len with bit j set, enumerated densely: m -> x
    const uint32_t period = 2u << j, rem = len & (period - 1);
    const uint32_t cnt = ((len >> (j + 1)) << j) + (rem > (1u << j) ? rem - (1u << j) : 0u);
    coop_group_sum(sm, 32, cnt, [](unsigned) { return true; }, [&](unsigned, uint32_t m) -> const g1_xyzz* {
        const uint32_t xi = ((m >> j) << (j + 1)) | (1u << j) | (m & ((1u << j) - 1));
        return arr + xi;
    });
    const unsigned dbl_n = is_row ? h + j : j;
    for (unsigned i = 0; i < dbl_n; i++) coop_dbl(sm, (threadIdx.x & 31) == 0);
    if ((threadIdx.x & 31) == 0) cs_emit(sm, AX, threadIdx.x >> 5, 0, planes + (size_t)pb * 32 + blockIdx.x);
}

// one block per polynomial: sum of its <= 32 weighted plane sums
__global__ void __launch_bounds__(128, 4) msm_final_coop_kernel(const g1_xyzz* planes, unsigned nplanes, g1_xyzz* out) {
    __shared__ CoopSm sm;
    planes += (size_t)blockIdx.x * 32;
    coop_group_sum(sm, 32, nplanes, [](unsigned) { return true; },
                   [&](unsigned, uint32_t m) -> const g1_xyzz* { return planes + m; });
    if ((threadIdx.x & 31) == 0) cs_emit(sm, AX, threadIdx.x >> 5, 0, out + blockIdx.x);
}

// T[w][i] = 2^c * T[w-1][i]: one thread per point walks the windows (load-time only).  The conversions to
// affine share inversions in groups of 8 windows (Montgomery's trick; the numerators wait in the table slot
// itself): W = 16 costs 2 Fermat chains per point instead of 15.  Same field elements as xyzz_to_affine.
// Row 0 is the SRS itself (96-byte points, kept for download / trim); the table has 128-byte entries.
__global__ void __launch_bounds__(128) srs_table_window_kernel(const g1_affine* powers, g1_tab* table, size_t n,
                                                              unsigned c, unsigned W) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr unsigned G = 8;
    g1_affine p = msm_ld_affine(powers + i);
    msm_st_tab(table + i, p);
    g1_xyzz acc = g1_xyzz::inf();
    xyzz_madd(acc, p);
    for (unsigned w0 = 1; w0 < W; w0 += G) {
        const unsigned cnt = W - w0 < G ? W - w0 : G;
        fq_t zz[G], zzz[G], pre[G];
        fq_t run = fq_t::one();
#pragma unroll 1
        for (unsigned g = 0; g < cnt; g++) {
#pragma unroll 1
            for (unsigned j = 0; j < c; j++) xyzz_dbl(acc);
            g1_affine xy;
            xy.x = acc.x; xy.y = acc.y;
            msm_st_tab(table + (size_t)(w0 + g) * n + i, xy);     // X, Y: divided below
            zz[g] = acc.zz;                                       // zero marks the point at infinity
            zzz[g] = acc.is_inf() ? fq_t::one() : acc.zzz;
            pre[g] = run;
            run = run * zzz[g];
        }
        fq_t inv = fq_inverse_bingcd(run);
#pragma unroll 1
        for (unsigned g = cnt; g-- > 0;) {
            const fq_t t = inv * pre[g];                          // 1 / ZZZ_g
            inv = inv * zzz[g];
            g1_tab* slot = table + (size_t)(w0 + g) * n + i;
            g1_affine out = g1_affine::inf();
            if (!zz[g].is_zero()) {
                const g1_affine xy = msm_ld_tab(slot);
                const fq_t zi = zz[g] * t;                        // 1 / ZZ = (ZZ / ZZZ)^2
                out.x = xy.x * sqr(zi);
                out.y = xy.y * t;
            }
            msm_st_tab(slot, out);
        }
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

// x = X / ZZ, y = Y / ZZZ through the binary-GCD inversion (88 multiplication times instead of the 618 of a
// Fermat chain): same field elements as xyzz_to_affine
__device__ __forceinline__ g1_affine xyzz_to_affine_gcd(const g1_xyzz& a) {
    if (a.is_inf()) return g1_affine::inf();
    const fq_t t = fq_inverse_bingcd(a.zzz);
    const fq_t zinv = a.zz * t;
    g1_affine r;
    r.x = a.x * sqr(zinv);
    r.y = a.y * t;
    return r;
}

// base[w] = 2^(8 w) G: one thread per window walks its doublings (the longest chain is 248 doublings)
__global__ void srs_table_bases_kernel(g1_xyzz* base) {
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= FB_W) return;
    const g1_affine g = g1_generator();
    g1_xyzz b = g1_xyzz::inf();
    xyzz_madd(b, g);
    for (unsigned i = 0; i < 8 * w; i++) xyzz_dbl(b);
    base[w] = b;
}

// tbl[w][d - 1] = d * base[w]: one thread per entry, 8-bit double-and-add and its own conversion to affine
// (was: one warp, 255 serial additions each followed by a Fermat chain -- 151 ms per SRS)
__global__ void __launch_bounds__(128) srs_table_kernel(const g1_xyzz* base, g1_affine* tbl) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= FB_W * FB_D) return;
    const unsigned w = t / FB_D, d = t % FB_D + 1;
    const g1_xyzz b = base[w];
    g1_xyzz acc = g1_xyzz::inf();
    for (int bit = 7; bit >= 0; bit--) {
        xyzz_dbl(acc);
        if ((d >> bit) & 1u) xyzz_add(acc, b);
    }
    tbl[t] = xyzz_to_affine_gcd(acc);
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
    out[i] = xyzz_to_affine_gcd(acc);
}

int srs_generate(zkp_ctx* ctx, const fr_t& tau, size_t first, size_t n, g1_affine* out_dev) {
    g1_affine* tbl = nullptr;
    g1_xyzz* base = nullptr;
    ZKP_CUDA(ctx, cudaMalloc(&tbl, sizeof(g1_affine) * FB_W * FB_D));
    ZKP_CUDA(ctx, cudaMalloc(&base, sizeof(g1_xyzz) * FB_W));
    srs_table_bases_kernel<<<1, 32, 0, ctx->stream>>>(base);
    ZKP_LAUNCHED(ctx);
    srs_table_kernel<<<(FB_W * FB_D + 127) / 128, 128, 0, ctx->stream>>>(base, tbl);
    ZKP_LAUNCHED(ctx);
    if (n) {
        srs_powers_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(tbl, tau, first, n, out_dev);
        ZKP_LAUNCHED(ctx);
    }
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZKP_CUDA(ctx, cudaFree(tbl));
    ZKP_CUDA(ctx, cudaFree(base));
    return ZKP_OK;
}

// ------------------------------------------------------------------ host driver
// acc <- acc + b on the host (add-2008-s / dbl-2008-s-1, same case analysis as g1.cuh): the G partial sums
// of a sharded commitment are combined here, G - 1 additions per commitment
static void host_xyzz_add(g1_xyzz& acc, const g1_xyzz& b) {
    using namespace hostfq;
    if (b.is_inf()) return;
    if (acc.is_inf()) { acc = b; return; }
    fq ax, ay, azz, azzz, bx, by, bzz, bzzz;
    memcpy(ax.l, acc.x.l, 48); memcpy(ay.l, acc.y.l, 48); memcpy(azz.l, acc.zz.l, 48); memcpy(azzz.l, acc.zzz.l, 48);
    memcpy(bx.l, b.x.l, 48); memcpy(by.l, b.y.l, 48); memcpy(bzz.l, b.zz.l, 48); memcpy(bzzz.l, b.zzz.l, 48);
    const fq U1 = mul(ax, bzz), S1 = mul(ay, bzzz);
    const fq Pp = sub(mul(bx, azz), U1), R = sub(mul(by, azzz), S1);
    fq X3, Y3, ZZ3, ZZZ3;
    if (is_zero(Pp)) {
        if (!is_zero(R) || is_zero(ay)) { acc = g1_xyzz::inf(); return; }
        const fq U = add(ay, ay), V = mul(U, U), W = mul(U, V), S = mul(ax, V), X2 = mul(ax, ax);
        const fq M = add(add(X2, X2), X2);
        X3 = sub(mul(M, M), add(S, S));
        Y3 = sub(mul(M, sub(S, X3)), mul(W, ay));
        ZZ3 = mul(V, azz); ZZZ3 = mul(W, azzz);
    } else {
        const fq PP = mul(Pp, Pp), PPP = mul(Pp, PP), Q = mul(U1, PP);
        X3 = sub(sub(mul(R, R), PPP), add(Q, Q));
        Y3 = sub(mul(R, sub(Q, X3)), mul(S1, PPP));
        ZZ3 = mul(mul(azz, bzz), PP); ZZZ3 = mul(mul(azzz, bzzz), PPP);
    }
    memcpy(acc.x.l, X3.l, 48); memcpy(acc.y.l, Y3.l, 48); memcpy(acc.zz.l, ZZ3.l, 48); memcpy(acc.zzz.l, ZZZ3.l, 48);
}
// x = X / ZZ, y = Y / ZZZ with 1/ZZ = (ZZ / ZZZ)^2.  The count <= MSM_MAX_BATCH points of one commit
// group share a single Fermat inversion (Montgomery's trick: prefix products, one inverse, unwind).
static void host_xyzz_to_affine_batch(const g1_xyzz* a, unsigned count, g1_affine* out) {
    hostfq::fq zzz[MSM_MAX_BATCH], pre[MSM_MAX_BATCH], acc;
    memcpy(acc.l, hostfq::ONE, 48);
    for (unsigned i = 0; i < count; i++) {
        out[i] = g1_affine::inf();
        if (a[i].is_inf()) continue;
        memcpy(zzz[i].l, a[i].zzz.l, 48);
        pre[i] = acc;                       // product of the finite ZZZ before i
        acc = hostfq::mul(acc, zzz[i]);
    }
    hostfq::fq inv = hostfq::inv(acc);      // 1 / (product of every finite ZZZ)
    for (unsigned i = count; i-- > 0;) {
        if (a[i].is_inf()) continue;
        const hostfq::fq t = hostfq::mul(inv, pre[i]);   // 1 / ZZZ_i
        inv = hostfq::mul(inv, zzz[i]);
        hostfq::fq X, Y, ZZ;
        memcpy(X.l, a[i].x.l, 48); memcpy(Y.l, a[i].y.l, 48); memcpy(ZZ.l, a[i].zz.l, 48);
        const hostfq::fq zi = hostfq::mul(ZZ, t);
        const hostfq::fq x = hostfq::mul(X, hostfq::mul(zi, zi)), y = hostfq::mul(Y, t);
        memcpy(out[i].x.l, x.l, 48); memcpy(out[i].y.l, y.l, 48);
    }
}

// Window width for an SRS of n powers, calibrated on B200 (bench/sweep_window.sh): the n * W mixed
// additions (10 Fq mul each) dominate; the sort / merge / bucket-reduction side is latency-bound and
// roughly constant up to 2^16 buckets, then grows ~160 Fq-mul-equivalents per extra bucket.
//   measured best: c = 16 for 2^16 .. 2^21 (W = 16), c = 20 from 2^22 (W = 13).
unsigned msm_choose_window(size_t n) {
    if (n == 0) n = 1;
    unsigned best = 4;
    double best_cost = 1e300;
    if (n < (1u << 14)) {
        for (unsigned c = 4; c <= 16; c++) {
            const unsigned W = 255 / c + 1;
            const double cost = 10.0 * (double)n * W + 56.0 * (double)(1u << (c - 1));
            if (cost < best_cost) { best_cost = cost; best = c; }
        }
        return best;
    }
    // only widths whose top window keeps >= 8 scalar bits (255 mod c): a short top window
    // funnels n / 2^bits entries into a handful of buckets (c = 17: the carry alone)
    static const unsigned cand[] = {13, 16, 20, 22};
    for (unsigned c : cand) {
        const unsigned W = 255 / c + 1;
        const double B = (double)(1u << (c - 1));
        const double cost = 10.0 * (double)n * W + 160.0 * (B > 65536.0 ? B - 65536.0 : 0.0);
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

// Fills rows 1 .. W-1 of the table from row 0 (the SRS powers themselves).
int srs_build_table(zkp_ctx* ctx, zkp_srs* srs) {
    if (srs->n == 0) return ZKP_OK;
    if (!srs->tab) {
        cudaError_t e = cudaMalloc(&srs->tab, (size_t)srs->W * srs->n * sizeof(g1_tab));
        if (e != cudaSuccess) {
            cuda_fail(ctx, e, "cudaMalloc(srs window table)", __FILE__, __LINE__);
            return e == cudaErrorMemoryAllocation ? ZKP_ERR_NOMEM : ZKP_ERR_CUDA;
        }
    }
    srs_table_window_kernel<<<(unsigned)((srs->n + 127) / 128), 128, 0, ctx->stream>>>(srs->d, srs->tab, srs->n, srs->c, srs->W);
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

static void msm_scratch_release(MsmScratch* s) {
    if (!s) return;
    cudaFree(s->digits); cudaFree(s->sorted); cudaFree(s->counts); cudaFree(s->offsets); cudaFree(s->cursor);
    cudaFree(s->giant); cudaFree(s->meta); cudaFree(s->tile_sums); cudaFree(s->buckets); cudaFree(s->slots); cudaFree(s->planes);
    cudaFree(s->top); cudaFree(s->totals); cudaFree(s->plan); cudaFree(s->aff[0]); cudaFree(s->aff[1]);
    delete s;
}

// built in a local and published to the context only when every allocation and occupancy query has
// succeeded: a half-initialised scratch must never be seen by the next commit
static int msm_scratch_build(zkp_ctx* ctx, MsmScratch* s) {
    ZKP_CUDA(ctx, cudaMalloc(&s->top, sizeof(long long)));
    ZKP_CUDA(ctx, cudaMalloc(&s->meta, 4 * MSM_MAX_BATCH * sizeof(uint32_t)));
    ZKP_CUDA(ctx, cudaMalloc(&s->tile_sums, (size_t)MSM_MAX_LEVELS * 1024 * MSM_MAX_BATCH * sizeof(uint32_t)));
    ZKP_CUDA(ctx, cudaMalloc(&s->totals, MSM_MAX_LEVELS * MSM_MAX_BATCH * sizeof(uint32_t)));
    int mb = 3;  // measured best on B200 (2^22: 22.3 / 21.8 / 23.1 / 23.7 / 24.4 ms for 2..6 blocks per SM)
    if (const char* e = getenv("ZKP_MSM_BLOCKS_PER_SM")) {  // tuning knob: 2 or 3, applies to every job size
        mb = atoi(e);
        s->acc_variant_forced = true;
    }
    mb = mb <= 2 ? 2 : 3;
    s->acc_variant = mb;
    int nb = 0;
    if (mb == 2) ZKP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, msm_accumulate_kernel<2, true>, 128, 0));
    else ZKP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, msm_accumulate_kernel<3, true>, 128, 0));
    s->acc_blocks_per_sm = nb > 0 ? nb : 1;
    int nb2 = 0;
    ZKP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, msm_accumulate_kernel<2, true>, 128, 0));
    s->acc_blocks_per_sm2 = nb2 > 0 ? nb2 : 1;
    int nb3 = 0;
    ZKP_CUDA(ctx, cudaFuncSetAttribute(msm_affine_round_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AFF_SMEM));
    ZKP_CUDA(ctx, cudaFuncSetAttribute(msm_affine_round_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AFF_SMEM));
    // three 60 KB blocks per SM need the large shared-memory carve-out; the default heuristic may pick less
    ZKP_CUDA(ctx, cudaFuncSetAttribute(msm_affine_round_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ZKP_CUDA(ctx, cudaFuncSetAttribute(msm_affine_round_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ZKP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb3, msm_affine_round_kernel<true>, 128, AFF_SMEM));
    s->aff_blocks_per_sm = nb3 > 0 ? nb3 : 1;
    if (const char* e = getenv("ZKP_MSM_AFFINE_ROUNDS")) s->aff_rounds_forced = atoi(e);
    if (const char* e = getenv("ZKP_MSM_AFFINE_MIN")) s->aff_min_pairs = (size_t)atoll(e);
    if (const char* e = getenv("ZKP_MSM_AFFINE_KMAX")) { s->aff_kmax = (size_t)atoll(e); if (s->aff_kmax < 8) s->aff_kmax = 8; }
    return ZKP_OK;
}

static int msm_scratch(zkp_ctx* ctx, MsmScratch** out) {
    if (!ctx->msm) {
        MsmScratch* s = new MsmScratch();
        const int rc = msm_scratch_build(ctx, s);
        if (rc) { msm_scratch_release(s); return rc; }
        ctx->msm = s;
    }
    *out = ctx->msm;
    return ZKP_OK;
}

void msm_free(zkp_ctx* ctx) {
    msm_scratch_release(ctx->msm);
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

// All-gather of every rank's slot (8 XYZZ partial sums + meta words), then the same host-side combination
// on every rank: G - 1 additions per commitment, one shared inversion, OR of the overflow flags.
static int msm_gather_partials(zkp_ctx* ctx, zkp_comm* cm, unsigned nb, g1_affine* out_host, int* overflow) {
    cudaStream_t st = ctx->stream;
    int rc;
    if ((rc = comm_allgather(cm, cm->gsend, cm->grecv, cm->slot_bytes, st))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(cm->hrecv, cm->grecv, cm->slot_bytes * (size_t)cm->nranks, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(ctx, cudaStreamSynchronize(st));
    g1_xyzz acc[MSM_MAX_BATCH];
    for (unsigned b = 0; b < nb; b++) { acc[b] = g1_xyzz::inf(); overflow[b] = 0; }
    for (int r = 0; r < cm->nranks; r++) {
        const uint8_t* slot = cm->hrecv + (size_t)r * cm->slot_bytes;
        const g1_xyzz* ps = reinterpret_cast<const g1_xyzz*>(slot);
        const uint32_t* pm = reinterpret_cast<const uint32_t*>(slot + 8 * sizeof(g1_xyzz));
        for (unsigned b = 0; b < nb; b++) {
            host_xyzz_add(acc[b], ps[b]);
            if (pm[4 * b + 2]) overflow[b] = 1;
        }
    }
    host_xyzz_to_affine_batch(acc, nb, out_host);
    return ZKP_OK;
}

// nb <= MSM_MAX_BATCH commitments against the same SRS in one set of launches.  lens[b] may exceed
// srs->n: coefficients beyond the SRS must be zero, otherwise overflow[b] = 1 (commit's Err) and
// out_host[b] is unspecified.
int msm_run_batch(zkp_ctx* ctx, const zkp_srs* srs, const fr_t* const* scalars_dev, const size_t* lens, unsigned nb,
                  g1_affine* out_host, int* overflow) {
    return msm_run_batch_ex(ctx, nullptr, srs, scalars_dev, lens, nullptr, nb, out_host, overflow);
}

// cm / offs: this rank's share of a commitment sharded over GPUs -- scalar i of polynomial b multiplies
// SRS power offs[b] + i; the ranks' partial sums (still XYZZ) and overflow flags are all-gathered and every
// rank finishes with the same affine commitments.  EVERY rank of the communicator must make the call, also
// with an empty range.
int msm_run_batch_ex(zkp_ctx* ctx, zkp_comm* cm, const zkp_srs* srs, const fr_t* const* scalars_dev, const size_t* lens,
                     const size_t* offs, unsigned nb, g1_affine* out_host, int* overflow) {
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (nb == 0 || nb > MSM_MAX_BATCH) return ZKP_ERR_INVALID;
    const bool shared = cm && cm->nranks > 1;
    const unsigned c = srs->c, W = srs->W;
    if ((size_t)W * srs->n >= (1ull << 31)) return ZKP_ERR_INVALID;
    MsmBatch batch;
    memset(&batch, 0, sizeof batch);
    size_t maxlen = 0, n = 0;  // n = scalars that can produce entries (<= srs->n)
    for (unsigned b = 0; b < nb; b++) {
        if (lens[b] >= (1ull << 31)) return ZKP_ERR_INVALID;
        const size_t off = offs ? offs[b] : 0;
        if (off >= (1ull << 31)) return ZKP_ERR_INVALID;
        batch.sc[b] = scalars_dev[b];
        batch.len[b] = (uint32_t)lens[b];
        batch.off[b] = (uint32_t)off;
        if (lens[b] > maxlen) maxlen = lens[b];
        const size_t room = off < srs->n ? srs->n - off : 0;
        const size_t eff = lens[b] < room ? lens[b] : room;
        if (eff > n) n = eff;
        ctx->msm_points += eff;
    }
    int local_ovf[MSM_MAX_BATCH] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool have_work = maxlen != 0 && n != 0;
    if (!have_work) {
        // nothing can be non-zero below the SRS length; still honour the overflow check
        for (unsigned b = 0; b < nb; b++) {
            long long top = -1;
            if (lens[b] && (rc = msm_highest_nonzero(ctx, scalars_dev[b], lens[b], &top))) return rc;
            local_ovf[b] = top >= 0;   // every scalar of this call lies beyond the SRS
            overflow[b] = local_ovf[b];
            memset(&out_host[b], 0, sizeof(g1_affine));
        }
        if (!shared) return ZKP_OK;
    }
    MsmScratch* s;
    if ((rc = msm_scratch(ctx, &s))) return rc;
    if (!have_work) {
        // an empty range still takes part in the gather: infinity partial sums + the flags found above
        cudaStream_t st0 = ctx->stream;
        uint32_t* hm0 = reinterpret_cast<uint32_t*>(ctx->pinned);
        memset(hm0, 0, 4 * MSM_MAX_BATCH * sizeof(uint32_t));
        for (unsigned b = 0; b < nb; b++) hm0[4 * b + 2] = (uint32_t)local_ovf[b];
        ZKP_CUDA(ctx, cudaMemsetAsync(cm->gsend, 0, cm->slot_bytes, st0));
        ZKP_CUDA(ctx, cudaMemcpyAsync(cm->gsend + 8 * sizeof(g1_xyzz), hm0, 4 * MSM_MAX_BATCH * sizeof(uint32_t),
                                      cudaMemcpyHostToDevice, st0));
        return msm_gather_partials(ctx, cm, nb, out_host, overflow);
    }

    const uint32_t B = 1u << (c - 1);
    const size_t E = (size_t)W * n;  // upper bound on the entries of one polynomial
    if (E >= (1ull << 31)) return ZKP_ERR_INVALID;
    // batched-affine pairing rounds while a round still has enough pairs to amortise its inversions
    unsigned R = 0;
    while (R + 1 < MSM_MAX_LEVELS && ((E * nb) >> (R + 1)) >= s->aff_min_pairs) R++;
    if (s->aff_rounds_forced >= 0) R = (unsigned)s->aff_rounds_forced < MSM_MAX_LEVELS - 1 ? (unsigned)s->aff_rounds_forced : MSM_MAX_LEVELS - 1;
    const unsigned levels = R + 1;
    const size_t ER = (E >> R) + (R ? B : 0);   // upper bound on what is left for the XYZZ pass
    // chunk length: one wave of resident threads when the job is small, 128-entry chunks (many
    // waves, negligible tail) when it is large
    // small jobs (about one wave) run best with the unconstrained 2-blocks/SM build, large ones with 3
    const int variant = (s->acc_variant_forced || ER * nb >= ((size_t)1 << 23)) ? s->acc_variant : 2;
    size_t resident =
        (size_t)ctx->sm_count * (variant == 2 ? s->acc_blocks_per_sm2 : s->acc_blocks_per_sm) * 128;
    if (resident == 0) resident = 128;
    size_t L = (ER * nb + resident - 1) / resident;
    if (L < 8) L = 8;
    if (L > 128) L = 128;
    const uint32_t nchunks = (uint32_t)((ER + L - 1) / L);
    // two-level bucket reduction: weight v = hi * 2^h + lo
    const unsigned h = c / 2;
    const uint32_t ncols = 1u << h, nrows = (B >> h) + 1;
    unsigned planes_r = 0, planes_c = h;
    while ((1u << planes_r) < nrows) planes_r++;   // bits of the largest row index (nrows - 1)
    if (planes_r + planes_c > 32) return ZKP_ERR_INVALID;
    const size_t nplanes = (size_t)nb * ((size_t)(nrows + ncols) + 32 + 1);

    if ((rc = ensure(ctx, &s->digits, &s->cap_entries, E * nb))) return rc;
    if ((rc = ensure(ctx, &s->sorted, &s->cap_sorted, E * nb))) return rc;
    if (s->cap_buckets < (size_t)B * nb) {
        const size_t need = (size_t)B * nb;
        size_t c1 = s->cap_buckets, c3 = c1, c4 = c1, c5 = c1;
        if ((rc = ensure(ctx, &s->counts, &c1, need))) return rc;
        if ((rc = ensure(ctx, &s->cursor, &c3, need))) return rc;
        if ((rc = ensure(ctx, &s->giant, &c4, need))) return rc;
        if ((rc = ensure(ctx, &s->buckets, &c5, need))) return rc;
        s->cap_buckets = need;
    }
    if ((rc = ensure(ctx, &s->offsets, &s->cap_offsets, (size_t)levels * B * nb))) return rc;
    if ((rc = ensure(ctx, &s->slots, &s->cap_slots, (size_t)2 * nchunks * nb))) return rc;
    if ((rc = ensure(ctx, &s->planes, &s->cap_planes, nplanes))) return rc;
    // outputs of round r (1-based) per polynomial: at most (E >> r) + B
    const size_t half = (E >> 1) + B, quarter = (E >> 2) + B;
    if (R >= 1) {
        if ((rc = ensure(ctx, &s->plan, &s->cap_plan, half * nb))) return rc;
        if ((rc = ensure(ctx, &s->aff[0], &s->cap_aff[0], half * nb))) return rc;
    }
    if (R >= 2 && (rc = ensure(ctx, &s->aff[1], &s->cap_aff[1], quarter * nb))) return rc;
    g1_xyzz* rowcol = s->planes;
    g1_xyzz* planes = rowcol + (size_t)nb * (nrows + ncols);
    g1_xyzz* sums = planes + (size_t)nb * 32;

    cudaStream_t st = ctx->stream;
    {
    ProfScope prof(ctx, "msm_sort");
    ZKP_CUDA(ctx, cudaMemsetAsync(s->counts, 0, (size_t)B * nb * sizeof(uint32_t), st));
    ZKP_CUDA(ctx, cudaMemsetAsync(s->meta, 0, 4 * MSM_MAX_BATCH * sizeof(uint32_t), st));
    msm_digits_kernel<<<dim3((unsigned)((maxlen + 255) / 256), nb), 256, 0, st>>>(
        batch, (uint32_t)n, (uint32_t)srs->n, c, W, s->digits, s->counts, s->meta);
    ZKP_LAUNCHED(ctx);
    {
        const uint32_t tiles = (B + SCAN_TILE - 1) / SCAN_TILE;  // <= 1024 for c <= 22
        msm_scan_tiles_kernel<<<dim3(tiles, levels * nb), SCAN_T, 0, st>>>(s->counts, B, tiles, nb, s->tile_sums);
        ZKP_LAUNCHED(ctx);
        msm_scan_sums_kernel<<<levels * nb, 1024, 0, st>>>(s->tile_sums, tiles, nb, s->totals, s->meta);
        ZKP_LAUNCHED(ctx);
        msm_scan_apply_kernel<<<dim3(tiles, levels * nb), SCAN_T, 0, st>>>(s->counts, B, tiles, nb, s->tile_sums,
                                                                           s->offsets, s->cursor);
        ZKP_LAUNCHED(ctx);
    }
    msm_scatter_kernel<<<dim3((unsigned)((n + 255) / 256), nb), 256, 0, st>>>(
        batch, s->digits, (uint32_t)n, W, (uint32_t)srs->n, B, s->cursor, s->sorted);
    ZKP_LAUNCHED(ctx);
    }
    {
    ProfScope prof(ctx, "msm_accumulate");
    // pairing rounds: level r - 1 -> level r; outputs ping-pong between aff[0] (odd rounds) and aff[1]
    for (unsigned r = 1; r <= R; r++) {
        const size_t bound = (E >> r) + B;                    // outputs of this round per polynomial
        // pairs per thread: whole waves of resident blocks (a partly filled last wave costs as much as a full
        // one), at most kmax pairs per inversion, at least 8
        const size_t slots = (size_t)ctx->sm_count * s->aff_blocks_per_sm;
        const size_t kmax = s->aff_kmax;
        size_t waves = (bound * nb + slots * 128 * kmax - 1) / (slots * 128 * kmax);
        if (waves == 0) waves = 1;
        size_t K = (bound * nb + waves * slots * 128 - 1) / (waves * slots * 128);
        if (K < 8) K = 8;
        if (K > kmax) K = kmax;
        msm_affine_plan_kernel<<<dim3((unsigned)((bound + 256 * PLAN_PER - 1) / (256 * PLAN_PER)), nb), 256, 0, st>>>(
            s->offsets, s->counts, s->totals, B, nb, r, half, s->plan);
        ZKP_LAUNCHED(ctx);
        const dim3 rgrid((unsigned)((bound + K * 128 - 1) / (K * 128)), nb);
        g1_affine* outp = s->aff[(r - 1) & 1];
        const size_t out_stride = ((r - 1) & 1) ? quarter : half;
        if (r == 1) {
            msm_affine_first_kernel<<<rgrid, 128, 0, st>>>(srs->tab, s->sorted, E, outp, out_stride, s->plan, half,
                                                           s->totals + r * MSM_MAX_BATCH, (uint32_t)K);
        } else {
            const g1_affine* inp = s->aff[r & 1];
            const size_t in_stride = (r & 1) ? quarter : half;
            msm_affine_round_kernel<false><<<rgrid, 128, AFF_SMEM, st>>>(nullptr, nullptr, 0, inp, in_stride, outp, out_stride,
                                                                  s->plan, half, s->totals + r * MSM_MAX_BATCH, (uint32_t)K);
        }
        ZKP_LAUNCHED(ctx);
    }
    const dim3 agrid((nchunks + 127) / 128, nb);
    const uint32_t* offR = s->offsets + (size_t)R * nb * B;
    const uint32_t* totR = s->totals + R * MSM_MAX_BATCH;
    if (R == 0) {
#define ZKP_ACC(...) msm_accumulate_kernel<__VA_ARGS__><<<agrid, 128, 0, st>>>( \
        srs->tab, s->sorted, offR, s->counts, totR, 0u, B, (uint32_t)L, nchunks, E, s->buckets, s->slots)
        if (variant == 2) ZKP_ACC(2, true);
        else ZKP_ACC(3, true);
#undef ZKP_ACC
    } else {
        const g1_affine* pts = s->aff[(R - 1) & 1];
        const size_t pstride = ((R - 1) & 1) ? quarter : half;
#define ZKP_ACC(...) msm_accumulate_kernel<__VA_ARGS__><<<agrid, 128, 0, st>>>( \
        pts, nullptr, offR, s->counts, totR, R, B, (uint32_t)L, nchunks, pstride, s->buckets, s->slots)
        if (variant == 2) ZKP_ACC(2, false);
        else ZKP_ACC(3, false);
#undef ZKP_ACC
    }
    ZKP_LAUNCHED(ctx);
    }
    if (ctx->after_accumulate) {
        void (*hook)(void*) = ctx->after_accumulate;
        ctx->after_accumulate = nullptr;
        hook(ctx->after_accumulate_arg);
    }
    {
    ProfScope prof(ctx, "msm_reduce");
    const uint32_t* offR = s->offsets + (size_t)R * nb * B;
    msm_merge_kernel<<<dim3((B + 127) / 128, nb), 128, 0, st>>>(offR, s->counts, R, B, (uint32_t)L, nchunks, s->slots,
                                                               s->buckets, s->giant, s->meta);
    ZKP_LAUNCHED(ctx);
    {
        // a giant bucket spans > GIANT_PARTS chunks, so there are fewer than nchunks / GIANT_PARTS
        const uint32_t gmax = nchunks / GIANT_PARTS < B ? nchunks / GIANT_PARTS : B;
        if (gmax) {
            msm_merge_giant_kernel<<<dim3(gmax, nb), 128, 128 * sizeof(g1_xyzz), st>>>(
                offR, s->counts, R, B, (uint32_t)L, nchunks, s->slots, s->buckets, s->giant, s->meta);
            ZKP_LAUNCHED(ctx);
        }
    }
    {
        // lanes per row / column sum: 32 while the blocks fit in one wave (shortest chain), fewer as the
        // launch outgrows the GPU (less idle tree work per sum)
        const size_t slots = (size_t)ctx->sm_count * 4, sums_total = (size_t)(nrows + ncols) * nb;
        unsigned lpo = 32;
        while (lpo > 4 && sums_total * lpo > slots * 32) lpo >>= 1;
        const unsigned R = 32 / lpo;
        msm_rowcol_coop_kernel<<<dim3((nrows + R - 1) / R + (ncols + R - 1) / R, nb), 128, 0, st>>>(
            s->buckets, B, h, nrows, ncols, lpo, rowcol);
        ZKP_LAUNCHED(ctx);
        msm_planes_coop_kernel<<<dim3(planes_r + planes_c, nb), 128, 0, st>>>(rowcol, h, nrows, ncols, planes_r, planes);
        ZKP_LAUNCHED(ctx);
        msm_final_coop_kernel<<<nb, 128, 0, st>>>(planes, planes_r + planes_c, sums);
        ZKP_LAUNCHED(ctx);
    }
    }
    if (shared) {
        ZKP_CUDA(ctx, cudaMemcpyAsync(cm->gsend, sums, nb * sizeof(g1_xyzz), cudaMemcpyDeviceToDevice, st));
        ZKP_CUDA(ctx, cudaMemcpyAsync(cm->gsend + 8 * sizeof(g1_xyzz), s->meta, 4 * MSM_MAX_BATCH * sizeof(uint32_t),
                                      cudaMemcpyDeviceToDevice, st));
        return msm_gather_partials(ctx, cm, nb, out_host, overflow);
    }
    // the single inversion of each conversion to affine runs on the host (one Fq Fermat chain
    // would occupy one GPU thread for ~0.6 ms)
    g1_xyzz* hs = reinterpret_cast<g1_xyzz*>(ctx->pinned);
    uint32_t* hm = reinterpret_cast<uint32_t*>(hs + MSM_MAX_BATCH);
    ZKP_CUDA(ctx, cudaMemcpyAsync(hs, sums, nb * sizeof(g1_xyzz), cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(ctx, cudaMemcpyAsync(hm, s->meta, 4 * MSM_MAX_BATCH * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(ctx, cudaStreamSynchronize(st));
    for (unsigned b = 0; b < nb; b++) overflow[b] = hm[4 * b + 2] != 0;
    host_xyzz_to_affine_batch(hs, nb, out_host);
    return ZKP_OK;
}

// A batch of commitments split over the ranks of `cm` by SRS ranges: the coefficients that can meet an SRS
// power, [0, min(len, srs->n)), are dealt out evenly; the last rank also takes whatever lies beyond the SRS
// (it must be zero: commit's degree check).  cm == nullptr or one rank: the plain batch.
int msm_commit_sharded(zkp_ctx* ctx, zkp_comm* cm, const zkp_srs* srs, const fr_t* const* polys, const size_t* lens,
                       unsigned nb, g1_affine* out_host, int* overflow) {
    if (!cm || cm->nranks == 1) return msm_run_batch(ctx, srs, polys, lens, nb, out_host, overflow);
    if (nb == 0 || nb > MSM_MAX_BATCH) return ZKP_ERR_INVALID;
    const fr_t* ptrs[MSM_MAX_BATCH];
    size_t ls[MSM_MAX_BATCH], offs[MSM_MAX_BATCH];
    const size_t G = (size_t)cm->nranks, r = (size_t)cm->rank;
    for (unsigned b = 0; b < nb; b++) {
        const size_t eff = lens[b] < srs->n ? lens[b] : srs->n;
        const size_t lo = eff * r / G;
        const size_t hi = (r + 1 == G) ? lens[b] : eff * (r + 1) / G;
        ptrs[b] = polys[b] + lo;
        ls[b] = hi - lo;
        offs[b] = lo;
    }
    return msm_run_batch_ex(ctx, cm, srs, ptrs, ls, offs, nb, out_host, overflow);
}

// msm_curve_addition over the first n powers (n <= srs->n checked by the caller).
int msm_run(zkp_ctx* ctx, const zkp_srs* srs, const fr_t* scalars_dev, size_t n, g1_affine* out_host) {
    if (n == 0) {
        memset(out_host, 0, sizeof(g1_affine));
        return ZKP_OK;
    }
    int ovf = 0;
    return msm_run_batch(ctx, srs, &scalars_dev, &n, 1, out_host, &ovf);
}

}  // namespace zkp

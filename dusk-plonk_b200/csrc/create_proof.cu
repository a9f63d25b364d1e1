// Native round driver: Prover::create_proof (src/prover.rs:67-474) behind one C-ABI call.
//
// The reference's prover is compiled host code that calls the NTT / commit primitives between
// Fiat-Shamir transcript operations.  This file is that host side above the kernels: the five
// rounds in the reference's order, the Merlin transcript (STROBE-128 over Keccak-f[1600]; labels
// src/prover.rs:99-105,139-199,203-226,268-295,321-405,435-450), the scalar side of the
// linearisation (src/prover/linearization_poly.rs:75-105,136-225) and the proof wire format
// (src/prover/proof.rs:36-66).  Polynomials stay in HBM for the whole proof; per proof the host
// reads back 11 points and 26 field elements at six synchronisation points (four commit groups, the
// permutation product, one batched opening of all 25 evaluations) and sends 8 challenges.
//
// Streams: the 8n-coset transforms of a, b, c, d, PI (and later z) depend on nothing the
// transcript still has to produce, so they run on a second stream, queued behind the accumulate
// kernel of the wire / z commitment (zkp_ctx::after_accumulate): they execute under its
// latency-bound bucket reduction instead of competing with the accumulation for the integer pipe.
// Event dependencies only, no host synchronisation beyond the reads the transcript needs.
#include <string.h>

#include <new>

#include "common.cuh"
#include "comm.cuh"
#include "host_driver.h"
#include "host_inv.h"

struct zkp_prover {
    zkp_ctx* ctx = nullptr;
    zkp_ctx* side = nullptr;          // second stream on the same device
    const zkp_srs* srs = nullptr;
    zkp_proving_key key;
    size_t n = 0, S = 0;
    unsigned k = 0;
    zkp_buf *W = nullptr, *Z = nullptr, *P7 = nullptr, *E7 = nullptr, *T = nullptr, *R = nullptr, *AGG = nullptr,
            *WZ = nullptr, *SAGG = nullptr, *WZW = nullptr;
    cudaEvent_t ev_main = nullptr, ev_side = nullptr;
    uint64_t n_inv[4] = {0, 0, 0, 0};   // 1 / n, Montgomery (host constant of the L1 polynomial)
    // circuit wiring for the witness gather on the device (zkp_prover_set_wiring)
    uint32_t* wire_idx = nullptr;   // [4][m]: witness index of wire j at gate i (src/prover.rs:114-119)
    uint32_t* pi_idx = nullptr;     // gate positions of the public inputs
    size_t m = 0, pi_count = 0, wv_cap = 0;
    // one proof over several GPUs (zkp_prover_create_sharded): this rank owns cosets u0 .. u0 + nloc - 1 of the
    // 8n domain (nl = nloc * n evaluations per vector) and coefficient slab [rank * slab, (rank + 1) * slab) of
    // every n-chunk of the quotient
    zkp_comm* comm = nullptr;
    bool sharded = false;
    unsigned nloc = 8, u0 = 0;
    size_t nl = 0, slab = 0;
    zkp_buf *TL = nullptr, *YA = nullptr, *TS = nullptr;   // local quotient values, exchanged terms, combined slabs
    bool wiring_set = false;        // zkp_prover_set_wiring has been called (prove_witness needs it)
    size_t max_wire_idx = 0;        // largest witness index the wiring reads (a witness array must cover it)
    zkp::fr_t* wv = nullptr;        // staging for the witness values and public-input values
};

namespace zkp {
namespace drv {

// wires[j][i] = witness[idx[j][i]] for i < m, zero in the padding rows (src/prover.rs:109-119)
__global__ void gather_wires_kernel(const fr_t* witness, size_t num_w, const uint32_t* idx, size_t m, size_t n,
                                    fr_t* wires) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 4 * n) return;
    const size_t j = t / n, i = t - j * n;
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (i < m) {
        const uint32_t k = idx[j * m + i];
        if (k < num_w) {
            const uint4* q = reinterpret_cast<const uint4*>(witness + k);
            a = __ldg(q); b = __ldg(q + 1);
        }
    }
    uint4* o = reinterpret_cast<uint4*>(wires + t);
    o[0] = a; o[1] = b;
}

// *flag = 1 if any of the n elements is non-zero (degree check of the quotient's upper chunks)
__global__ void any_nonzero_kernel(const fr_t* p, size_t n, uint32_t* flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* q = reinterpret_cast<const uint4*>(p + i);
    const uint4 a = q[0], b = q[1];
    if (a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) *flag = 1u;
}

// dense public-input vector (src/lib.rs:206-219): pi[idx[c]] = values[c] on a zeroed n-vector
__global__ void scatter_pi_kernel(const fr_t* values, const uint32_t* idx, size_t count, size_t n, fr_t* pi) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= count) return;
    const uint32_t i = idx[c];
    if (i < n) pi[i] = values[c];
}

// t_(i' + n q) = sum_u c[q][u] Y_u[i'], c[q][u] = g^(-n q) w_8^(-u q): the cross-coset step of the 8n-point
// inverse coset transform (derivation: ntt.cu, "cosets of the n-domain").  y[u][i'], out[q][i'], i' < slab.
struct CombineArgs { fr_t c[8][8]; const fr_t* y; fr_t* out; size_t slab; };

__global__ void __launch_bounds__(128) coset8_combine_kernel(const __grid_constant__ CombineArgs a) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.slab) return;
    fr_t y[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const uint4* p = reinterpret_cast<const uint4*>(a.y + (size_t)u * a.slab + i);
        const uint4 lo = p[0], hi = p[1];
        y[u].l[0] = lo.x; y[u].l[1] = lo.y; y[u].l[2] = lo.z; y[u].l[3] = lo.w;
        y[u].l[4] = hi.x; y[u].l[5] = hi.y; y[u].l[6] = hi.z; y[u].l[7] = hi.w;
    }
#pragma unroll 1
    for (int q = 0; q < 8; q++) {
        fr_t t = a.c[q][0] * y[0];
#pragma unroll
        for (int u = 1; u < 8; u++) t = t + a.c[q][u] * y[u];
        uint4* o = reinterpret_cast<uint4*>(a.out + (size_t)q * a.slab + i);
        o[0] = make_uint4(t.l[0], t.l[1], t.l[2], t.l[3]);
        o[1] = make_uint4(t.l[4], t.l[5], t.l[6], t.l[7]);
    }
}

static int coset8_combine(zkp_ctx* ctx, const fr_t* y, fr_t* out, size_t slab, unsigned k) {
    CombineArgs a;
    const fr w8i = F::inv(fr_load(reinterpret_cast<const uint64_t*>(fft_constant_host(3, 0).l)));
    const fr gni = F::inv(F::pow(F::from_u64(7), (uint64_t)1 << k));
    fr gq = F::one();
    for (int q = 0; q < 8; q++) {
        const fr wq = F::pow(w8i, (uint64_t)q);     // w_8^-q
        fr v = gq;                                   // g^(-n q) w_8^(-q u)
        for (int u = 0; u < 8; u++) {
            fr_store(reinterpret_cast<uint64_t*>(a.c[q][u].l), v);
            v = F::mul(v, wq);
        }
        gq = F::mul(gq, gni);
    }
    a.y = y; a.out = out; a.slab = slab;
    ProfScope prof(ctx, "ntt");
    coset8_combine_kernel<<<(unsigned)((slab + 127) / 128), 128, 0, ctx->stream>>>(a);
    ZKP_LAUNCHED(ctx);
    return ZKP_OK;
}

#define TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

static inline zkp_poly_ref ref(const zkp_buf* b, size_t off, size_t len) { zkp_poly_ref r; r.buf = b; r.off = off; r.len = len; return r; }

// key.poly / key.eval8 indices
enum { Q_M = 0, Q_L, Q_R, Q_O, Q_C, Q_D, Q_ARITH, Q_RANGE, Q_LOGIC, Q_FIXED, Q_VAR, S1, S2, S3, S4 };

// Transcript-independent coset transforms, queued on the second stream from inside the commit (see
// zkp_ctx::after_accumulate): they start when the commit's accumulate kernel has finished and run under
// its bucket reduction.
struct SideJob {
    zkp_prover* pr;
    const fr_t* in;
    size_t in_stride, len_in;
    fr_t* out;
    unsigned batch;
    int rc;
};

static void run_side_job(void* arg) {
    SideJob* j = static_cast<SideJob*>(arg);
    zkp_prover* pr = j->pr;
    zkp_ctx *ctx = pr->ctx, *side = pr->side;
    const size_t n8 = 8 * pr->n;
    j->rc = ZKP_ERR_CUDA;
    if (cudaEventRecord(pr->ev_main, ctx->stream) != cudaSuccess) return;
    if (cudaStreamWaitEvent(side->stream, pr->ev_main, 0) != cudaSuccess) return;
    if (pr->sharded) {   // the rank's cosets of each polynomial: n-point transforms, no data from other ranks
        j->rc = ZKP_OK;
        for (unsigned b = 0; b < j->batch && j->rc == ZKP_OK; b++)
            j->rc = coset8_forward(side, j->in + b * j->in_stride, j->len_in, j->out + b * pr->nl, pr->k, pr->u0, pr->nloc);
    } else {
        j->rc = ntt_run(side, j->in, j->in_stride, j->len_in, j->out, n8, pr->k + 3, false, true, j->batch);
    }
    if (j->rc == ZKP_OK && cudaEventRecord(pr->ev_side, side->stream) != cudaSuccess) j->rc = ZKP_ERR_CUDA;
}

// commit with a side job attached: if the MSM never reached its accumulate kernel (all-zero input),
// the job is queued right after
static int commit_group(zkp_prover* pr, const zkp_poly_ref* polys, unsigned count, uint64_t* out_xy);
static int commit_group_with(zkp_prover* pr, const zkp_poly_ref* polys, unsigned count, uint64_t* out_xy, SideJob* job) {
    pr->ctx->after_accumulate = run_side_job;
    pr->ctx->after_accumulate_arg = job;
    job->rc = ZKP_OK;
    const int rc = commit_group(pr, polys, count, out_xy);
    if (pr->ctx->after_accumulate) {   // not consumed
        pr->ctx->after_accumulate = nullptr;
        run_side_job(job);
    }
    if (rc) return rc;
    return job->rc;
}

static int commit_group(zkp_prover* pr, const zkp_poly_ref* polys, unsigned count, uint64_t* out_xy) {
    const fr_t* ptrs[8];
    size_t lens[8];
    int ovf[8];
    for (unsigned i = 0; i < count; i++) { ptrs[i] = polys[i].buf->d + polys[i].off; lens[i] = polys[i].len; }
    TRY(msm_commit_sharded(pr->ctx, pr->comm, pr->srs, ptrs, lens, count, reinterpret_cast<g1_affine*>(out_xy), ovf));
    for (unsigned i = 0; i < count; i++) if (ovf[i]) return ZKP_ERR_DEGREE;  // commit(..)? in the reference
    return ZKP_OK;
}

// Rounds 4 / 5 of a proof shared by several GPUs, on the slabs where the quotient's inverse transform left them:
// rank s works on coefficients [lo, hi) = [s slab, (s + 1) slab) of every polynomial (the last rank also on
// the few coefficients at and beyond n).  Openings are sums of per-slab Horner values weighed by z^lo (one
// gather of the partial values); the linearisation and
// the aggregated witness polynomials are built slab by slab; the division by (X - z) is a suffix scan whose
// carry into a slab is the Horner value of everything above it (one gather of two values).  Same field
// elements, same transcript, same commitments as on one GPU.
static int commit_group(zkp_prover* pr, const zkp_poly_ref* polys, unsigned count, uint64_t* out_xy);
static int openings_sharded(zkp_prover* pr, Transcript& tr, const fr ch[7], const fr& zc, const fr& zw,
                            const zkp_poly_ref wp[4], const zkp_poly_ref& zp, uint64_t* comms, uint64_t* evals,
                            uint8_t* proof_bytes, uint8_t* transcript_out) {
    zkp_ctx* ctx = pr->ctx;
    zkp_comm* cm = pr->comm;
    const zkp_proving_key& key = pr->key;
    const size_t n = pr->n, slab = pr->slab;
    const unsigned G = cm ? (unsigned)cm->nranks : 1u, rank = cm ? (unsigned)cm->rank : 0u;
    const bool last = rank + 1 == G;
    const size_t lo = (size_t)rank * slab;
    cudaStream_t st = ctx->stream;
    // a rank's range of a replicated polynomial of `len` coefficients
    auto part = [&](const zkp_poly_ref& p) {
        const size_t hi = last ? p.len : (lo + slab < p.len ? lo + slab : p.len);
        return ref(p.buf, p.off + (lo < p.len ? lo : p.len), hi > lo ? hi - lo : 0);
    };
    // ---- openings: 8 chunk slabs of t, 20 polynomials at z, 4 at z w
    enum { E_T0 = 0, E_A = 4, E_B, E_C, E_D, E_S1, E_S2, E_S3, E_QARITH, E_QC, E_QL, E_QR,
           E_QM, E_QO, E_QD, E_QRANGE, E_QLOGIC, E_QFIXED, E_QVAR, E_Z, E_S4,
           E_AN, E_BN, E_DN, E_PERM, E_COUNT };
    static_assert(E_COUNT == 28, "one evaluation launch");
    const size_t tail = pr->srs->n - n;      // coefficients of t_4 at and beyond n: behind the last rank's chunk-3 slab
    zkp_poly_ref ep[E_COUNT];
    uint8_t which[E_COUNT];
    memset(which, 0, sizeof which);
    for (unsigned q = 0; q < 4; q++) ep[E_T0 + q] = ref(pr->TS, q * slab, slab + (q == 3 && last ? tail : 0));
    for (unsigned j = 0; j < 4; j++) ep[E_A + j] = part(wp[j]);
    ep[E_S1] = part(key.poly[S1]); ep[E_S2] = part(key.poly[S2]); ep[E_S3] = part(key.poly[S3]);
    ep[E_QARITH] = part(key.poly[Q_ARITH]); ep[E_QC] = part(key.poly[Q_C]); ep[E_QL] = part(key.poly[Q_L]);
    ep[E_QR] = part(key.poly[Q_R]); ep[E_QM] = part(key.poly[Q_M]); ep[E_QO] = part(key.poly[Q_O]);
    ep[E_QD] = part(key.poly[Q_D]); ep[E_QRANGE] = part(key.poly[Q_RANGE]); ep[E_QLOGIC] = part(key.poly[Q_LOGIC]);
    ep[E_QFIXED] = part(key.poly[Q_FIXED]); ep[E_QVAR] = part(key.poly[Q_VAR]); ep[E_Z] = part(zp);
    ep[E_S4] = part(key.poly[S4]);
    ep[E_AN] = part(wp[0]); ep[E_BN] = part(wp[1]); ep[E_DN] = part(wp[3]); ep[E_PERM] = part(zp);
    for (unsigned j = E_AN; j < E_COUNT; j++) which[j] = 1;
    uint64_t pts[8];
    memcpy(pts, zc.l, 32);
    memcpy(pts + 4, zw.l, 32);
    fr_t* res = nullptr;
    TRY(poly_eval2_launch(ctx, ep, which, E_COUNT, pts, &res));
    // gather the partial values
    const size_t blob = E_COUNT * sizeof(fr_t);
    uint8_t* send = cm ? cm->gsend : reinterpret_cast<uint8_t*>(pr->YA->d);
    uint8_t* recv = cm ? cm->grecv : reinterpret_cast<uint8_t*>(pr->YA->d + 64);
    ZKP_CUDA(ctx, cudaMemcpyAsync(send, res, E_COUNT * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
    TRY(comm_allgather(cm, send, recv, blob, st));
    uint8_t* hb = cm ? cm->hrecv : reinterpret_cast<uint8_t*>(ctx->pinned);
    ZKP_CUDA(ctx, cudaMemcpyAsync(hb, recv, blob * G, cudaMemcpyDeviceToHost, st));
    ZKP_CUDA(ctx, cudaStreamSynchronize(st));
    const fr z_n = F::pow(zc, (uint64_t)n);
    const fr zs = F::pow(zc, (uint64_t)slab), zws = F::pow(zw, (uint64_t)slab);
    fr full[E_COUNT];
    for (unsigned j = 0; j < E_COUNT; j++) full[j] = F::zero();
    {
        fr wz = F::one(), wzw = F::one();            // z^lo_s, (z w)^lo_s
        for (unsigned s2 = 0; s2 < G; s2++) {
            const uint64_t* pv = reinterpret_cast<const uint64_t*>(hb + blob * s2);
            for (unsigned j = 0; j < E_COUNT; j++)
                full[j] = F::add(full[j], F::mul(fr_load(pv + 4 * j), which[j] ? wzw : wz));
            wz = F::mul(wz, zs); wzw = F::mul(wzw, zws);
        }
    }
    fr t_eval = F::zero();
    {
        fr zq = F::one();
        for (unsigned q = 0; q < 4; q++) { t_eval = F::add(t_eval, F::mul(zq, full[E_T0 + q])); zq = F::mul(zq, z_n); }
    }
    auto E = [&](int i) { return full[i]; };
    const fr a = E(E_A), b = E(E_B), c = E(E_C), d = E(E_D);
    const fr s1 = E(E_S1), s2 = E(E_S2), s3 = E(E_S3);
    const fr qarith = E(E_QARITH), qc = E(E_QC), ql = E(E_QL), qr = E(E_QR);
    const fr an = E(E_AN), bn = E(E_BN), dn = E(E_DN), pe = E(E_PERM);
    fr sc[12];
    {
        fr ch8[8];
        for (unsigned j = 0; j < 7; j++) ch8[j] = ch[j];
        ch8[7] = zc;
        const fr e15[15] = {a, b, c, d, an, bn, dn, s1, s2, s3, qarith, qc, ql, qr, pe};
        linearization_scalars((uint64_t)n, ch8, e15, sc);
    }
    fr r_eval = F::zero();
    {
        static const int lin_eval[12] = {E_QM, E_QL, E_QR, E_QO, E_QD, E_QC, E_QRANGE, E_QLOGIC, E_QFIXED, E_QVAR, E_Z, E_S4};
        for (unsigned j = 0; j < 12; j++) r_eval = F::add(r_eval, F::mul(sc[j], E(lin_eval[j])));
    }
    const fr ev[16] = {a, b, c, d, an, bn, dn, s1, s2, s3, qarith, qc, ql, qr, pe, r_eval};
    static const char* const el[15] = {"a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval",
                                       "d_next_eval", "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval",
                                       "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "perm_eval"};
    for (unsigned j = 0; j < 15; j++) tr.append_scalar(el[j], ev[j]);
    tr.append_scalar("t_eval", t_eval);
    tr.append_scalar("r_eval", r_eval);

    // ---- r(X), the aggregated polynomials and their quotients by (X - z), (X - z w): this rank's range only
    const size_t agg_len = pr->srs->n > n + 3 ? pr->srs->n : n + 3;     // <= n + 8 here
    const size_t hiR = last ? n + 3 : lo + slab, hiA = last ? agg_len : lo + slab, hiS = last ? n + 3 : lo + slab;
    uint64_t scl[16 * 4];
    {
        static const int lin_key[10] = {Q_M, Q_L, Q_R, Q_O, Q_D, Q_C, Q_RANGE, Q_LOGIC, Q_FIXED, Q_VAR};
        zkp_poly_ref lr[12];
        for (unsigned j = 0; j < 10; j++) lr[j] = part(key.poly[lin_key[j]]);
        lr[10] = part(zp);
        lr[11] = part(key.poly[S4]);
        for (unsigned j = 0; j < 12; j++) fr_store(scl + 4 * j, sc[j]);
        TRY(zkp_poly_lincomb_dev(ctx, lr, scl, 12, pr->R, lo, hiR - lo));
    }
    const fr v1 = tr.challenge_scalar("v_challenge");
    const fr v2 = tr.challenge_scalar("v_challenge");     // nothing is appended in between (src/prover.rs:435-450)
    {
        // main part [lo, min(hi, n)): the four chunk slabs + the replicated polynomials' slabs
        zkp_poly_ref ar[12] = {ref(pr->TS, 0, slab), ref(pr->TS, slab, slab), ref(pr->TS, 2 * slab, slab),
                               ref(pr->TS, 3 * slab, slab), ref(pr->R, lo, slab), ref(wp[0].buf, wp[0].off + lo, slab),
                               ref(wp[1].buf, wp[1].off + lo, slab), ref(wp[2].buf, wp[2].off + lo, slab),
                               ref(wp[3].buf, wp[3].off + lo, slab), ref(key.poly[S1].buf, key.poly[S1].off + lo, slab),
                               ref(key.poly[S2].buf, key.poly[S2].off + lo, slab),
                               ref(key.poly[S3].buf, key.poly[S3].off + lo, slab)};
        fr as[12];
        as[0] = F::one(); as[1] = z_n; as[2] = F::sqr(z_n); as[3] = F::mul(as[2], z_n); as[4] = v1;
        for (unsigned j = 5; j < 12; j++) as[j] = F::mul(as[j - 1], v1);
        for (unsigned j = 0; j < 12; j++) fr_store(scl + 4 * j, as[j]);
        TRY(zkp_poly_lincomb_dev(ctx, ar, scl, 12, pr->AGG, lo, slab));
        if (last) {
            // tail [n, agg_len): z^3n * (t_4's coefficients beyond n, behind the chunk-3 slab) + v r + v^2 a + .. + v^5 d
            zkp_poly_ref tr6[6] = {ref(pr->TS, 4 * slab, tail), ref(pr->R, n, 3), ref(wp[0].buf, wp[0].off + n, 2),
                                   ref(wp[1].buf, wp[1].off + n, 2), ref(wp[2].buf, wp[2].off + n, 2),
                                   ref(wp[3].buf, wp[3].off + n, 2)};
            const fr ts[6] = {as[3], as[4], as[5], as[6], as[7], as[8]};
            for (unsigned j = 0; j < 6; j++) fr_store(scl + 4 * j, ts[j]);
            TRY(zkp_poly_lincomb_dev(ctx, tr6, scl, 6, pr->AGG, n, agg_len - n));
        }
        const zkp_poly_ref br[4] = {part(zp), part(wp[0]), part(wp[1]), part(wp[3])};
        fr bs[4];
        bs[0] = F::one();
        for (unsigned j = 1; j < 4; j++) bs[j] = F::mul(bs[j - 1], v2);
        for (unsigned j = 0; j < 4; j++) fr_store(scl + 4 * j, bs[j]);
        TRY(zkp_poly_lincomb_dev(ctx, br, scl, 4, pr->SAGG, lo, hiS - lo));
    }
    // Horner value of each rank's range -> the carries of the two divisions
    fr carry1 = F::zero(), carry2 = F::zero();
    if (G > 1) {
        const zkp_poly_ref hp[2] = {ref(pr->AGG, lo, hiA - lo), ref(pr->SAGG, lo, hiS - lo)};
        const uint8_t hw[2] = {0, 1};
        TRY(poly_eval2_launch(ctx, hp, hw, 2, pts, &res));
        ZKP_CUDA(ctx, cudaMemcpyAsync(send, res, 2 * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
        TRY(comm_allgather(cm, send, recv, 2 * sizeof(fr_t), st));
        ZKP_CUDA(ctx, cudaMemcpyAsync(hb, recv, 2 * sizeof(fr_t) * G, cudaMemcpyDeviceToHost, st));
        ZKP_CUDA(ctx, cudaStreamSynchronize(st));
        // c_s = P_(s+1) + z^(len_(s+1)) c_(s+1), from the top; len of the last range differs between the two
        const fr zlast1 = F::pow(zc, (uint64_t)(agg_len - (size_t)(G - 1) * slab));
        const fr zlast2 = F::pow(zw, (uint64_t)(n + 3 - (size_t)(G - 1) * slab));
        for (unsigned s2 = G - 1; s2 > rank; s2--) {
            const uint64_t* pv = reinterpret_cast<const uint64_t*>(hb + 2 * sizeof(fr_t) * s2);
            carry1 = F::add(fr_load(pv), F::mul(s2 == G - 1 ? zlast1 : zs, carry1));
            carry2 = F::add(fr_load(pv + 4), F::mul(s2 == G - 1 ? zlast2 : zws, carry2));
        }
    }
    {
        // [0, a_(lo+1) .. a_(hi-1), carry] / (X - z) -> w_lo .. w_(hi-1)   (the last rank has no carry element)
        uint64_t* hp2 = reinterpret_cast<uint64_t*>(ctx->pinned);
        memset(hp2, 0, 32);
        fr_store(hp2 + 4, carry1);
        fr_store(hp2 + 8, carry2);
        ZKP_CUDA(ctx, cudaMemcpyAsync(pr->AGG->d + lo, hp2, sizeof(fr_t), cudaMemcpyHostToDevice, st));
        ZKP_CUDA(ctx, cudaMemcpyAsync(pr->SAGG->d + lo, hp2, sizeof(fr_t), cudaMemcpyHostToDevice, st));
        if (!last) {
            ZKP_CUDA(ctx, cudaMemcpyAsync(pr->AGG->d + hiA, hp2 + 4, sizeof(fr_t), cudaMemcpyHostToDevice, st));
            ZKP_CUDA(ctx, cudaMemcpyAsync(pr->SAGG->d + hiS, hp2 + 8, sizeof(fr_t), cudaMemcpyHostToDevice, st));
        }
        const size_t l1 = hiA - lo + (last ? 0 : 1), l2 = hiS - lo + (last ? 0 : 1);
        TRY(zkp_poly_div_linear_dev(ctx, ref(pr->AGG, lo, l1), zc.l, pr->WZ, lo));
        TRY(zkp_poly_div_linear_dev(ctx, ref(pr->SAGG, lo, l2), zw.l, pr->WZW, lo));
        const fr_t* ptrs[2] = {pr->WZ->d + lo, pr->WZW->d + lo};
        const size_t lens[2] = {l1 - 1, l2 - 1}, offs[2] = {lo, lo};
        int ovf[2];
        TRY(msm_run_batch_ex(ctx, cm, pr->srs, ptrs, lens, offs, 2, reinterpret_cast<g1_affine*>(comms + 12 * 9), ovf));
        if (ovf[0] || ovf[1]) return ZKP_ERR_DEGREE;
    }
    for (unsigned j = 0; j < 16; j++) fr_store(evals + 4 * j, ev[j]);
    if (proof_bytes) {
        for (unsigned j = 0; j < 11; j++) g1_compress(comms + 12 * j, proof_bytes + 48 * j);
        static const int wire_order[16] = {0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 7, 8, 9, 15, 14};
        for (unsigned j = 0; j < 16; j++) fr_bytes(ev[wire_order[j]], proof_bytes + 48 * 11 + 32 * j);
    }
    if (transcript_out) tr.save(transcript_out);
    return ZKP_OK;
}

static int prove(zkp_prover* pr, const uint8_t transcript_in[203], const uint64_t* wires_host, const zkp_buf* wires_dev,
                 const uint64_t* pi_host, const zkp_buf* pi_dev, const uint64_t* blinders, uint64_t* comms,
                 uint64_t* evals, uint8_t* proof_bytes, uint8_t* transcript_out) {
    zkp_ctx* ctx = pr->ctx;
    const zkp_proving_key& key = pr->key;
    const size_t n = pr->n, S = pr->S, n8 = 8 * n;
    const unsigned k = pr->k, k8 = k + 3;
    TRY(set_device(ctx));
    Transcript tr;
    tr.load(transcript_in);

    // round 1: wires -> iNTT -> blind -> commit (src/prover.rs:107-158)
    nvtxRangePushA("create_proof");
    struct PopAll { int depth = 1; ~PopAll() { while (depth-- > 0) nvtxRangePop(); } } nvtx_guard;   // also on early return
    nvtxRangePushA("round 1: wire polynomials"); nvtx_guard.depth++;
    const zkp_buf* W = wires_dev;
    if (!W) {
        ZKP_CUDA(ctx, cudaMemcpyAsync(pr->W->d, wires_host, 4 * n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        W = pr->W;
    }
    TRY(ntt_run(ctx, W->d, n, n, pr->P7->d, S, k, true, false, 4));
    for (unsigned j = 0; j < 4; j++) TRY(zkp_poly_blind_dev(ctx, pr->P7, j * S, n, blinders + 8 * j, 2));
    if (pi_dev) {
        TRY(ntt_run(ctx, pi_dev->d, 0, n, pr->P7->d + 4 * S, 0, k, true, false, 1));
    } else if (!pi_host) {   // the dense vector was built in place (zkp_prover_prove_witness)
        TRY(ntt_run(ctx, pr->P7->d + 4 * S, 0, n, pr->P7->d + 4 * S, 0, k, true, false, 1));
    } else {
        ZKP_CUDA(ctx, cudaMemcpyAsync(pr->P7->d + 4 * S, pi_host, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        TRY(ntt_run(ctx, pr->P7->d + 4 * S, 0, n, pr->P7->d + 4 * S, 0, k, true, false, 1));
    }
    // a, b, c, d, PI on the 8n coset: second stream, under the bucket reduction of the wire commitments
    // (reference order: src/prover.rs:229, quotient_poly.rs:54-58,145 -- same values, earlier)
    zkp_poly_ref wp[4];
    for (unsigned j = 0; j < 4; j++) wp[j] = ref(pr->P7, j * S, n + 2);
    SideJob wires8 = {pr, pr->P7->d, S, n + 3, pr->E7->d, 5, ZKP_OK};   // sharded: local cosets, stride nl
    TRY(commit_group_with(pr, wp, 4, comms, &wires8));
    static const char* const wl[4] = {"a_w", "b_w", "c_w", "d_w"};
    for (unsigned j = 0; j < 4; j++) tr.append_commitment(wl[j], comms + 12 * j);

    // round 2: permutation accumulator (src/prover.rs:160-199)
    nvtxRangePop(); nvtxRangePushA("round 2: permutation z");
    const fr beta = tr.challenge_scalar("beta");
    tr.append_scalar("beta", beta);
    const fr gamma = tr.challenge_scalar("gamma");
    zkp_poly_ref wr[4];
    for (unsigned j = 0; j < 4; j++) wr[j] = ref(W, j * n, n);
    TRY(zkp_perm_z_dev(ctx, n, wr, key.sigma_evals, key.roots, beta.l, gamma.l, pr->Z, 0));
    const zkp_poly_ref zp = ref(pr->P7, 5 * S, n + 3);
    TRY(ntt_run(ctx, pr->Z->d, 0, n, pr->P7->d + 5 * S, 0, k, true, false, 1));
    TRY(zkp_poly_blind_dev(ctx, pr->P7, 5 * S, n, blinders + 32, 3));
    // z on the 8n coset needs no further challenge either: second stream, under the z commitment's reduction
    const size_t evn = pr->sharded ? pr->nl : n8;      // evaluations each coset-domain vector holds on this rank
    SideJob z8 = {pr, pr->P7->d + 5 * S, 0, n + 3, pr->E7->d + 5 * evn, 1, ZKP_OK};
    TRY(commit_group_with(pr, &zp, 1, comms + 12 * 4, &z8));
    tr.append_commitment("z", comms + 12 * 4);

    // round 3: quotient on the 8n coset (src/prover.rs:201-287, quotient_poly.rs)
    nvtxRangePop(); nvtxRangePushA("round 3: quotient");
    bool slab_openings = false;
    fr ch[7];
    ch[0] = tr.challenge_scalar("alpha");
    ch[1] = beta;
    ch[2] = gamma;
    ch[3] = tr.challenge_scalar("range separation challenge");
    ch[4] = tr.challenge_scalar("logic separation challenge");
    ch[5] = tr.challenge_scalar("fixed base separation challenge");
    ch[6] = tr.challenge_scalar("variable base separation challenge");
    const fr alpha = ch[0];
    const fr alpha2 = F::sqr(alpha);
    // L1 * alpha^2: idft of (alpha^2, 0, ..) has every coefficient alpha^2 / n (quotient_poly.rs:264-272)
    const fr n_inv = fr_load(pr->n_inv);
    const fr l1c = F::mul(alpha2, n_inv);
    TRY(zkp_buf_fill(ctx, pr->P7, 6 * S, n, l1c.l));
    if (pr->sharded) TRY(coset8_forward(ctx, pr->P7->d + 6 * S, n, pr->E7->d + 6 * evn, k, pr->u0, pr->nloc));
    else TRY(ntt_run(ctx, pr->P7->d + 6 * S, 0, n, pr->E7->d + 6 * n8, 0, k8, false, true, 1));
    ZKP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pr->ev_side, 0));
    zkp_quotient_args qa;
    memset(&qa, 0, sizeof qa);
    for (unsigned j = 0; j < 4; j++) { qa.wires[j] = ref(pr->E7, j * evn, evn); qa.sigma[j] = key.eval8[S1 + j]; }
    qa.pi = ref(pr->E7, 4 * evn, evn);
    qa.z = ref(pr->E7, 5 * evn, evn);
    qa.l1 = ref(pr->E7, 6 * evn, evn);
    for (unsigned j = 0; j < 11; j++) qa.sel[j] = key.eval8[j];
    qa.linear = key.linear8;
    for (unsigned j = 0; j < 7; j++) memcpy(qa.challenges[j], ch[j].l, 32);
    memcpy(qa.zh_inv, key.zh_inv, sizeof qa.zh_inv);
    qa.widget_mask = key.widget_mask;
    if (pr->sharded) {
        // quotient on this rank's cosets; then the 8n-point inverse coset transform as: n-point inverses of the
        // local cosets (scaled by h_u^-e / 8), ONE exchange (slab s of every coset to rank s), and the 8 x 8
        // combination across cosets, which leaves this rank with slab `rank` of every n-chunk of t(X);
        // all-gathered chunk by chunk into the natural coefficient order for rounds 4 / 5
        qa.coset_log_n = k;
        qa.coset_first = pr->u0;
        TRY(zkp_quotient_range_dev(ctx, 0, &qa, 0, pr->nl, pr->TL, 0));
        TRY(coset8_inverse_local(ctx, pr->TL->d, k, pr->u0, pr->nloc));
        TRY(comm_exchange_slabs(pr->comm, pr->TL->d, pr->YA->d, n, pr->nloc, ctx->stream));
        TRY(coset8_combine(ctx, pr->YA->d, pr->TS->d, pr->slab, k));
        // rounds 4 / 5 work on the slabs where they lie unless the SRS is long enough for t_4 to reach beyond its
        // first eight tail coefficients (then: gather t(X) chunk by chunk and continue as one GPU would)
        // (and the slab must hold those eight: tiny circuits on many GPUs take the gather path as well)
        // (and the local vectors are large enough to double as staging when there is no communicator)
        slab_openings = pr->srs->n <= n + 8 && pr->slab >= 8 && pr->nl >= 256;
        if (!slab_openings)
            for (unsigned q = 0; q < 8; q++)
                TRY(comm_allgather(pr->comm, pr->TS->d + q * pr->slab, pr->T->d + q * n, pr->slab * sizeof(fr_t), ctx->stream));
    } else {
        TRY(zkp_quotient_dev(ctx, k8, &qa, pr->T, 0));
        TRY(ntt_run(ctx, pr->T->d, 0, n8, pr->T->d, 0, k8, true, true, 1));  // coset_idft -> t coefficients
    }
    const zkp_poly_ref tq[4] = {ref(pr->T, 0, n), ref(pr->T, n, n), ref(pr->T, 2 * n, n), ref(pr->T, 3 * n, 5 * n)};
    if (slab_openings) {
        // This rank holds coefficients [lo, lo + slab) of every n-chunk of t.  t_4 = chunk 3 followed by the few
        // coefficients of chunk 4 that still meet an SRS power (they live on rank 0 and move to the last rank,
        // behind its chunk-3 slab, where the SRS continues); everything else of chunks 4 .. 7 must be zero --
        // commit's degree check -- which every rank verifies on its own slabs.  One small gather carries both.
        zkp_comm* cm = pr->comm;
        const unsigned G = cm ? (unsigned)cm->nranks : 1u, rank = cm ? (unsigned)cm->rank : 0u;
        const size_t slab = pr->slab, lo = (size_t)rank * slab, tail = pr->srs->n - n;   // <= 8
        const bool last = rank + 1 == G;
        uint8_t* send = cm ? cm->gsend : reinterpret_cast<uint8_t*>(pr->YA->d);
        uint8_t* recv = cm ? cm->grecv : reinterpret_cast<uint8_t*>(pr->YA->d + 64);
        uint8_t* hb = cm ? cm->hrecv : reinterpret_cast<uint8_t*>(ctx->pinned);
        const size_t blob = 9 * sizeof(fr_t);             // [8 head coefficients of chunk 4][flag word]
        ZKP_CUDA(ctx, cudaMemsetAsync(send, 0, blob, ctx->stream));
        ZKP_CUDA(ctx, cudaMemcpyAsync(send, pr->TS->d + 4 * slab, 8 * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
        const size_t skip = rank == 0 ? tail : 0;
        any_nonzero_kernel<<<(unsigned)((4 * slab - skip + 255) / 256), 256, 0, ctx->stream>>>(
            pr->TS->d + 4 * slab + skip, 4 * slab - skip, reinterpret_cast<uint32_t*>(send + 8 * sizeof(fr_t)));
        ZKP_LAUNCHED(ctx);
        TRY(comm_allgather(cm, send, recv, blob, ctx->stream));
        ZKP_CUDA(ctx, cudaMemcpyAsync(hb, recv, blob * G, cudaMemcpyDeviceToHost, ctx->stream));
        if (last && G > 1)   // rank 0's head of chunk 4 -> behind this rank's chunk-3 slab (G == 1: already there)
            ZKP_CUDA(ctx, cudaMemcpyAsync(pr->TS->d + 4 * slab, recv, 8 * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
        ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (unsigned r2 = 0; r2 < G; r2++)
            if (*reinterpret_cast<const uint32_t*>(hb + blob * r2 + 8 * sizeof(fr_t))) return ZKP_ERR_DEGREE;
        const fr_t* ptrs[4];
        size_t lens[4], offs[4];
        int ovf[4];
        for (unsigned q = 0; q < 4; q++) { ptrs[q] = pr->TS->d + q * slab; lens[q] = slab; offs[q] = lo; }
        if (last) lens[3] = slab + tail;
        TRY(msm_run_batch_ex(ctx, cm, pr->srs, ptrs, lens, offs, 4, reinterpret_cast<g1_affine*>(comms + 12 * 5), ovf));
        for (unsigned q = 0; q < 4; q++) if (ovf[q]) return ZKP_ERR_DEGREE;
    } else {
        TRY(commit_group(pr, tq, 4, comms + 12 * 5));
    }
    static const char* const tl[4] = {"t_low", "t_mid", "t_high", "t_4"};
    for (unsigned j = 0; j < 4; j++) tr.append_commitment(tl[j], comms + 12 * (5 + j));

    // rounds 4/5: evaluations, linearisation, openings (src/prover.rs:289-452)
    nvtxRangePop(); nvtxRangePushA("rounds 4-5: openings");
    const fr zc = tr.challenge_scalar("z_challenge");
    const fr zw = F::mul(zc, fr_load(key.generator));
    if (slab_openings)
        return openings_sharded(pr, tr, ch, zc, zw, wp, zp, comms, evals, proof_bytes, transcript_out);
    // Every opening of the proof in ONE batched launch and one read-back: the 12 evaluations at z, the 4
    // at z w, and the nine linearisation polynomials not among them at z -- r(X) = sum_j sc_j p_j(X) is
    // linear in the p_j, so r(z) = sum_j sc_j p_j(z) needs no evaluation of r itself
    // (linearization_poly.rs:52-73,107-110: same field element).
    enum { E_T = 0, E_A, E_B, E_C, E_D, E_S1, E_S2, E_S3, E_QARITH, E_QC, E_QL, E_QR,       // at z
           E_QM, E_QO, E_QD, E_QRANGE, E_QLOGIC, E_QFIXED, E_QVAR, E_Z, E_S4,               // at z (for r)
           E_AN, E_BN, E_DN, E_PERM, E_COUNT };                                             // at z w
    zkp_poly_ref ep[E_COUNT];
    uint8_t which[E_COUNT];
    memset(which, 0, sizeof which);
    // the t_4 commitment above passed its degree check, so every coefficient of t_4 = T[3n ..] at or beyond the
    // SRS length is zero: the quotient is opened, aggregated and divided over its real length only (same
    // field elements as over the zero-padded 8n / 5n vectors of quotient_poly.rs:115, prover.rs:253-259,422-434)
    const size_t t4_len = pr->srs->n < 5 * n ? pr->srs->n : 5 * n;
    const size_t agg_len = t4_len > n + 3 ? t4_len : n + 3;
    ep[E_T] = ref(pr->T, 0, 3 * n + t4_len);
    for (unsigned j = 0; j < 4; j++) ep[E_A + j] = wp[j];
    ep[E_S1] = key.poly[S1]; ep[E_S2] = key.poly[S2]; ep[E_S3] = key.poly[S3];
    ep[E_QARITH] = key.poly[Q_ARITH]; ep[E_QC] = key.poly[Q_C]; ep[E_QL] = key.poly[Q_L]; ep[E_QR] = key.poly[Q_R];
    ep[E_QM] = key.poly[Q_M]; ep[E_QO] = key.poly[Q_O]; ep[E_QD] = key.poly[Q_D];
    ep[E_QRANGE] = key.poly[Q_RANGE]; ep[E_QLOGIC] = key.poly[Q_LOGIC]; ep[E_QFIXED] = key.poly[Q_FIXED];
    ep[E_QVAR] = key.poly[Q_VAR]; ep[E_Z] = zp; ep[E_S4] = key.poly[S4];
    ep[E_AN] = wp[0]; ep[E_BN] = wp[1]; ep[E_DN] = wp[3]; ep[E_PERM] = zp;
    for (unsigned j = E_AN; j < E_COUNT; j++) which[j] = 1;
    uint64_t pts[8], ev_raw[E_COUNT * 4];
    memcpy(pts, zc.l, 32);
    memcpy(pts + 4, zw.l, 32);
    TRY(zkp_poly_eval2_dev(ctx, ep, which, E_COUNT, pts, ev_raw));
    auto E = [&](int i) { return fr_load(ev_raw + 4 * i); };
    const fr t_eval = E(E_T);
    const fr a = E(E_A), b = E(E_B), c = E(E_C), d = E(E_D);
    const fr s1 = E(E_S1), s2 = E(E_S2), s3 = E(E_S3);
    const fr qarith = E(E_QARITH), qc = E(E_QC), ql = E(E_QL), qr = E(E_QR);
    const fr an = E(E_AN), bn = E(E_BN), dn = E(E_DN), pe = E(E_PERM);

    // r(X) = sum scalar * poly(X): the polynomial itself is still needed for the opening witness
    fr sc[12];
    zkp_poly_ref lr[12];
    static const int lin_key[10] = {Q_M, Q_L, Q_R, Q_O, Q_D, Q_C, Q_RANGE, Q_LOGIC, Q_FIXED, Q_VAR};
    for (unsigned j = 0; j < 10; j++) lr[j] = key.poly[lin_key[j]];
    lr[10] = zp;
    lr[11] = key.poly[S4];
    {
        fr ch8[8];
        for (unsigned j = 0; j < 7; j++) ch8[j] = ch[j];
        ch8[7] = zc;
        const fr e15[15] = {a, b, c, d, an, bn, dn, s1, s2, s3, qarith, qc, ql, qr, pe};
        linearization_scalars((uint64_t)n, ch8, e15, sc);
    }
    uint64_t scl[16 * 4];
    for (unsigned j = 0; j < 12; j++) fr_store(scl + 4 * j, sc[j]);
    TRY(zkp_poly_lincomb_dev(ctx, lr, scl, 12, pr->R, 0, n + 3));
    const zkp_poly_ref rr = ref(pr->R, 0, n + 3);
    fr r_eval = F::zero();
    {
        static const int lin_eval[12] = {E_QM, E_QL, E_QR, E_QO, E_QD, E_QC, E_QRANGE, E_QLOGIC, E_QFIXED, E_QVAR, E_Z, E_S4};
        for (unsigned j = 0; j < 12; j++) r_eval = F::add(r_eval, F::mul(sc[j], E(lin_eval[j])));
    }

    // evaluations in `Evaluations` order (prover.py EVAL_NAMES)
    const fr ev[16] = {a, b, c, d, an, bn, dn, s1, s2, s3, qarith, qc, ql, qr, pe, r_eval};
    static const char* const el[15] = {"a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval",
                                       "d_next_eval", "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval",
                                       "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "perm_eval"};
    for (unsigned j = 0; j < 15; j++) tr.append_scalar(el[j], ev[j]);
    tr.append_scalar("t_eval", t_eval);
    tr.append_scalar("r_eval", r_eval);

    // W_z: (t_low + z^n t_mid + z^2n t_high + z^3n t_4) + v r + v^2 a + .., divided by (X - z)
    {
        const fr z_n = F::pow(zc, (uint64_t)n);
        const fr v1 = tr.challenge_scalar("v_challenge");
        zkp_poly_ref ar[12] = {tq[0], tq[1], tq[2], ref(pr->T, 3 * n, t4_len), rr, wp[0], wp[1], wp[2], wp[3],
                               key.poly[S1], key.poly[S2], key.poly[S3]};
        fr as[12];
        as[0] = F::one();
        as[1] = z_n;
        as[2] = F::sqr(z_n);
        as[3] = F::mul(as[2], z_n);
        as[4] = v1;
        for (unsigned j = 5; j < 12; j++) as[j] = F::mul(as[j - 1], v1);
        for (unsigned j = 0; j < 12; j++) fr_store(scl + 4 * j, as[j]);
        TRY(zkp_poly_lincomb_dev(ctx, ar, scl, 12, pr->AGG, 0, agg_len));
        TRY(zkp_poly_div_linear_dev(ctx, ref(pr->AGG, 0, agg_len), zc.l, pr->WZ, 0));
        // the second v_challenge follows the first with nothing appended in between
        // (src/prover.rs:435-450): both witnesses are committed as one batch of two
        const fr v2 = tr.challenge_scalar("v_challenge");
        const zkp_poly_ref br[4] = {zp, wp[0], wp[1], wp[3]};
        fr bs[4];
        bs[0] = F::one();
        for (unsigned j = 1; j < 4; j++) bs[j] = F::mul(bs[j - 1], v2);
        for (unsigned j = 0; j < 4; j++) fr_store(scl + 4 * j, bs[j]);
        TRY(zkp_poly_lincomb_dev(ctx, br, scl, 4, pr->SAGG, 0, n + 3));
        TRY(zkp_poly_div_linear_dev(ctx, ref(pr->SAGG, 0, n + 3), zw.l, pr->WZW, 0));
        const zkp_poly_ref wc[2] = {ref(pr->WZ, 0, agg_len - 1), ref(pr->WZW, 0, n + 2)};
        TRY(commit_group(pr, wc, 2, comms + 12 * 9));
    }

    for (unsigned j = 0; j < 16; j++) fr_store(evals + 4 * j, ev[j]);
    if (proof_bytes) {
        // 11 compressed G1 then the 16 evaluations in the field order of `Evaluations`
        // (src/prover/linearization_poly.rs:113-130): a b c d a' b' d' q_arith q_c q_l q_r s1 s2 s3 r perm
        for (unsigned j = 0; j < 11; j++) g1_compress(comms + 12 * j, proof_bytes + 48 * j);
        static const int wire_order[16] = {0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 7, 8, 9, 15, 14};
        for (unsigned j = 0; j < 16; j++) fr_bytes(ev[wire_order[j]], proof_bytes + 48 * 11 + 32 * j);
    }
    if (transcript_out) tr.save(transcript_out);
    return ZKP_OK;
}

}  // namespace drv
}  // namespace zkp

using namespace zkp;

extern "C" {

static int drv_prover_create(zkp_ctx* ctx, zkp_comm* comm, bool sharded, const zkp_srs* srs, const zkp_proving_key* key,
                             zkp_prover** out);

int zkp_prover_destroy(zkp_prover* pr) {
    if (!pr) return ZKP_OK;
    zkp_ctx* ctx = pr->ctx;
    if (ctx) cudaSetDevice(ctx->device);
    if (pr->side) cudaStreamSynchronize(pr->side->stream);
    zkp_buf** bufs[] = {&pr->W, &pr->Z, &pr->P7, &pr->E7, &pr->T, &pr->R, &pr->AGG, &pr->WZ, &pr->SAGG, &pr->WZW,
                        &pr->TL, &pr->YA, &pr->TS};
    for (zkp_buf** b : bufs) if (*b) { zkp_buf_free(ctx, *b); *b = nullptr; }
    if (pr->wire_idx) cudaFree(pr->wire_idx);
    if (pr->pi_idx) cudaFree(pr->pi_idx);
    if (pr->wv) cudaFree(pr->wv);
    if (pr->ev_main) cudaEventDestroy(pr->ev_main);
    if (pr->ev_side) cudaEventDestroy(pr->ev_side);
    if (pr->side) zkp_ctx_destroy(pr->side);
    delete pr;
    return ZKP_OK;
}

int zkp_prover_create(zkp_ctx* ctx, const zkp_srs* srs, const zkp_proving_key* key, zkp_prover** out) {
    return drv_prover_create(ctx, nullptr, false, srs, key, out);
}

int zkp_prover_create_sharded(zkp_ctx* ctx, zkp_comm* comm, const zkp_srs* srs, const zkp_proving_key* key,
                              zkp_prover** out) {
    if (comm && comm->ctx != ctx) return ZKP_ERR_INVALID;
    return drv_prover_create(ctx, comm, true, srs, key, out);
}

static int drv_prover_create(zkp_ctx* ctx, zkp_comm* comm, bool sharded, const zkp_srs* srs, const zkp_proving_key* key,
                             zkp_prover** out) {
    if (!ctx || !srs || !key || !out || key->k < 1 || key->k > 25 || !key->roots) return ZKP_ERR_INVALID;
    const size_t n = (size_t)1 << key->k;
    const unsigned G = comm ? (unsigned)comm->nranks : 1u, rank = comm ? (unsigned)comm->rank : 0u;
    if (sharded && n < 8) return ZKP_ERR_INVALID;
    // evaluations each coset-domain vector of the key holds on this rank: all 8n, or the rank's 8 / G cosets
    const size_t n8 = sharded ? 8 * n / G : 8 * n;
    for (int j = 0; j < 15; j++) {
        const zkp_poly_ref &p = key->poly[j], &e = key->eval8[j];
        if (!p.buf || p.off + p.len > p.buf->n || p.len < n) return ZKP_ERR_INVALID;
        if (!e.buf || e.off + e.len > e.buf->n || e.len < n8) return ZKP_ERR_INVALID;
    }
    for (int j = 0; j < 4; j++) {
        const zkp_poly_ref& s = key->sigma_evals[j];
        if (!s.buf || s.off + s.len > s.buf->n || s.len < n) return ZKP_ERR_INVALID;
    }
    if (!key->linear8.buf || key->linear8.off + key->linear8.len > key->linear8.buf->n || key->linear8.len < n8 ||
        key->roots->n < n)
        return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    zkp_prover* pr = new (std::nothrow) zkp_prover();
    if (!pr) return ZKP_ERR_NOMEM;
    pr->ctx = ctx;
    pr->srs = srs;
    pr->key = *key;
    pr->n = n;
    pr->k = key->k;
    pr->S = n + 8;  // the seven polynomials bound for the 8n coset sit side by side
    pr->comm = comm;
    pr->sharded = sharded;
    if (sharded) {
        pr->nloc = 8 / G; pr->u0 = rank * pr->nloc;
        pr->nl = (size_t)pr->nloc * n; pr->slab = n / G;
    }
    drv::fr_store(pr->n_inv, drv::F::inv(drv::F::from_u64((uint64_t)n)));
    struct { zkp_buf** b; size_t len; } want[] = {
        {&pr->W, 4 * n}, {&pr->Z, n}, {&pr->P7, 7 * pr->S}, {&pr->E7, 7 * n8}, {&pr->T, 8 * n}, {&pr->R, n + 3},
        {&pr->AGG, 5 * n}, {&pr->WZ, 5 * n}, {&pr->SAGG, n + 3}, {&pr->WZW, n + 3},
        {&pr->TL, sharded ? n8 : 0}, {&pr->YA, sharded ? n8 : 0}, {&pr->TS, sharded ? n8 : 0}};
    for (auto& w : want)
        if (w.len && (rc = zkp_buf_alloc(ctx, w.len, w.b))) { zkp_prover_destroy(pr); return rc; }
    if ((rc = zkp_buf_zero(ctx, pr->P7, 0, 7 * pr->S)) || (rc = zkp_ctx_create(ctx->device, &pr->side))) {
        zkp_prover_destroy(pr);
        return rc;
    }
    if (cudaEventCreateWithFlags(&pr->ev_main, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pr->ev_side, cudaEventDisableTiming) != cudaSuccess) {
        zkp_prover_destroy(pr);
        return ZKP_ERR_CUDA;
    }
    if ((rc = zkp_ctx_sync(ctx))) { zkp_prover_destroy(pr); return rc; }
    *out = pr;
    return ZKP_OK;
}

int zkp_prover_prove(zkp_prover* pr, const uint8_t transcript[203], const uint64_t* wires_host,
                     const zkp_buf* wires_dev, const uint64_t* pi_host, const zkp_buf* pi_dev,
                     const uint64_t blinders[44], uint64_t commitments[132], uint64_t evaluations[64],
                     uint8_t proof_bytes[1040], uint8_t transcript_out[203]) {
    if (!pr || !transcript || !blinders || !commitments || !evaluations) return ZKP_ERR_INVALID;
    if ((!wires_host && !wires_dev) || (!pi_host && !pi_dev)) return ZKP_ERR_INVALID;
    if ((wires_dev && wires_dev->n < 4 * pr->n) || (pi_dev && pi_dev->n < pr->n)) return ZKP_ERR_INVALID;
    const int rc = drv::prove(pr, transcript, wires_host, wires_dev, pi_host, pi_dev, blinders, commitments,
                              evaluations, proof_bytes, transcript_out);
    if (rc) {
        // leave both streams idle so the next proof never races a half-finished one
        cudaStreamSynchronize(pr->side->stream);
        cudaStreamSynchronize(pr->ctx->stream);
    }
    return rc;
}

int zkp_prover_set_wiring(zkp_prover* pr, const uint32_t* wire_idx, size_t m, const uint32_t* pi_idx,
                          size_t pi_count) {
    if (!pr || (!wire_idx && m) || m > pr->n || (!pi_idx && pi_count) || pi_count > pr->n) return ZKP_ERR_INVALID;
    zkp_ctx* ctx = pr->ctx;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (pr->wire_idx) cudaFree(pr->wire_idx);
    if (pr->pi_idx) cudaFree(pr->pi_idx);
    pr->wire_idx = pr->pi_idx = nullptr;
    pr->m = pr->pi_count = 0;
    // a stale or foreign wiring must be an error, not a silently different circuit: public inputs sit on
    // gates of the n-domain, and prove_witness checks the witness array against the largest index read
    for (size_t i = 0; i < pi_count; i++) if (pi_idx[i] >= pr->n) return ZKP_ERR_INVALID;
    size_t mx = 0;
    for (size_t i = 0; i < 4 * m; i++) if (wire_idx[i] > mx) mx = wire_idx[i];
    pr->max_wire_idx = mx;
    if (m) {
        ZKP_CUDA(ctx, cudaMalloc(&pr->wire_idx, 4 * m * sizeof(uint32_t)));
        ZKP_CUDA(ctx, cudaMemcpy(pr->wire_idx, wire_idx, 4 * m * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (pi_count) {
        ZKP_CUDA(ctx, cudaMalloc(&pr->pi_idx, pi_count * sizeof(uint32_t)));
        ZKP_CUDA(ctx, cudaMemcpy(pr->pi_idx, pi_idx, pi_count * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    pr->m = m;
    pr->pi_count = pi_count;
    pr->wiring_set = true;
    return ZKP_OK;
}

int zkp_prover_prove_witness(zkp_prover* pr, const uint8_t transcript[203], const uint64_t* witness, size_t num_w,
                             const uint64_t* pi_values, const uint64_t blinders[44], uint64_t commitments[132],
                             uint64_t evaluations[64], uint8_t proof_bytes[1040], uint8_t transcript_out[203]) {
    if (!pr || !transcript || !blinders || !commitments || !evaluations || (!witness && num_w) ||
        (!pi_values && pr->pi_count) || (pr->m && num_w <= pr->max_wire_idx))
        return ZKP_ERR_INVALID;
    if (!pr->wiring_set) return ZKP_ERR_STATE;    // zkp_prover_set_wiring comes first
    zkp_ctx* ctx = pr->ctx;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    // a sharded proof: every rank has the same host witness, so each uploads 1 / G of it over its own PCIe
    // link and the ranks all-gather the slices over NVLink (the public-input values sit behind the padded witness)
    const size_t G = pr->sharded && pr->comm ? (size_t)pr->comm->nranks : 1;
    const size_t chunk = (num_w + G - 1) / G, wslots = chunk * G;
    const size_t n = pr->n, need = wslots + pr->pi_count + 1;
    if (pr->wv_cap < need) {
        ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (pr->wv) cudaFree(pr->wv);
        pr->wv = nullptr; pr->wv_cap = 0;
        ZKP_CUDA(ctx, cudaMalloc(&pr->wv, need * sizeof(fr_t)));
        pr->wv_cap = need;
    }
    cudaStream_t st = ctx->stream;
    if (G > 1 && num_w) {
        const size_t lo = (size_t)pr->comm->rank * chunk, hi = lo + chunk < num_w ? lo + chunk : num_w;
        if (hi > lo)
            ZKP_CUDA(ctx, cudaMemcpyAsync(pr->wv + lo, witness + 4 * lo, (hi - lo) * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        if ((rc = comm_allgather(pr->comm, pr->wv + lo, pr->wv, chunk * sizeof(fr_t), st))) return rc;   // in place
    } else if (num_w) {
        ZKP_CUDA(ctx, cudaMemcpyAsync(pr->wv, witness, num_w * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    }
    drv::gather_wires_kernel<<<(unsigned)((4 * n + 255) / 256), 256, 0, st>>>(pr->wv, num_w, pr->wire_idx, pr->m, n, pr->W->d);
    ZKP_LAUNCHED(ctx);
    fr_t* pi = pr->P7->d + 4 * pr->S;
    ZKP_CUDA(ctx, cudaMemsetAsync(pi, 0, n * sizeof(fr_t), st));
    if (pr->pi_count) {
        fr_t* pv = pr->wv + wslots;
        ZKP_CUDA(ctx, cudaMemcpyAsync(pv, pi_values, pr->pi_count * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        drv::scatter_pi_kernel<<<(unsigned)((pr->pi_count + 255) / 256), 256, 0, st>>>(pv, pr->pi_idx, pr->pi_count, n, pi);
        ZKP_LAUNCHED(ctx);
    }
    rc = drv::prove(pr, transcript, nullptr, pr->W, nullptr, nullptr, blinders, commitments, evaluations, proof_bytes,
                    transcript_out);
    if (rc) {
        cudaStreamSynchronize(pr->side->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    return rc;
}

/* Merlin transcript operations on a serialized state (host code): what TranscriptProtocol needs
 * outside a proof (seeding with the verification key, public inputs) without leaving native code. */
int zkp_transcript_init(uint8_t state[203], const uint8_t* label, uint32_t len) {
    if (!state || (!label && len)) return ZKP_ERR_INVALID;
    // STROBE-128 initial state: F([1, R + 2, 1, 0, 1, 96] || "STROBEv1.0.2"), then meta-AD of the
    // protocol label "Merlin v1.0" and Merlin's own domain separator carrying the caller's label
    drv::Transcript t;
    memset(t.st, 0, 200);
    const uint8_t hdr[6] = {1, drv::Transcript::RATE + 2, 1, 0, 1, 96};
    memcpy(t.st, hdr, 6);
    memcpy(t.st + 6, "STROBEv1.0.2", 12);
    uint64_t w[25];
    memcpy(w, t.st, 200);
    zkp_keccak_f1600(w);
    memcpy(t.st, w, 200);
    t.pos = t.pos_begin = t.cur_flags = 0;
    t.meta_ad("Merlin v1.0", 11, false);
    t.append_message("dom-sep", label, len);
    t.save(state);
    return ZKP_OK;
}

int zkp_transcript_append(uint8_t state[203], const char* label, const uint8_t* msg, uint32_t len) {
    if (!state || !label || (!msg && len)) return ZKP_ERR_INVALID;
    drv::Transcript t;
    t.load(state);
    t.append_message(label, msg, len);
    t.save(state);
    return ZKP_OK;
}

int zkp_transcript_challenge(uint8_t state[203], const char* label, uint8_t* out, uint32_t len) {
    if (!state || !label || (!out && len)) return ZKP_ERR_INVALID;
    drv::Transcript t;
    t.load(state);
    t.challenge_bytes(label, out, len);
    t.save(state);
    return ZKP_OK;
}

/* The scalar side of the linearisation alone (host code; unit-tested on the CPU against widgets.py):
 * challenges = alpha beta gamma range logic fixed var z (8 x 4, Montgomery), evals in `Evaluations`
 * order (first 15 used), out = the 12 scalars of q_m q_l q_r q_o q_4 q_c q_range q_logic q_fixed q_var z s_sigma_4. */
int zkp_linearization_scalars(unsigned k, const uint64_t challenges[32], const uint64_t evals[60], uint64_t out[48]) {
    if (!challenges || !evals || !out || k > 28) return ZKP_ERR_INVALID;
    using drv::fr; using drv::fr_load; using drv::fr_store;
    fr ch[8], e[15], sc[12];
    for (int j = 0; j < 8; j++) ch[j] = fr_load(challenges + 4 * j);
    for (int j = 0; j < 15; j++) e[j] = fr_load(evals + 4 * j);
    drv::linearization_scalars((uint64_t)1 << k, ch, e, sc);
    for (int j = 0; j < 12; j++) fr_store(out + 4 * j, sc[j]);
    return ZKP_OK;
}

/* Encodings the transcript and the proof wire format use (host code). */
int zkp_g1_compress(const uint64_t xy[12], uint8_t out[48]) {
    if (!xy || !out) return ZKP_ERR_INVALID;
    drv::g1_compress(xy, out);
    return ZKP_OK;
}

int zkp_fr_from_wide(const uint8_t bytes[64], uint64_t out_mont[4]) {
    if (!bytes || !out_mont) return ZKP_ERR_INVALID;
    drv::fr_store(out_mont, drv::fr_from_wide(bytes));
    return ZKP_OK;
}

}  // extern "C"

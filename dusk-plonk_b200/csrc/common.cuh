// Shared runtime state behind the C ABI (include/zkp_b200.h): context, device buffers,
// error plumbing and the launch counter bench.py reports as gpu_launches.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges are no-ops unless a profiler injects itself
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/zkp_b200.h"
#include "arith.cuh"
#include "g1.cuh"

namespace zkp {

struct NttDomain;  // ntt.cu
struct Coset8Tab;  // ntt.cu
struct TwTab;      // ntt.cu
struct MsmScratch; // msm.cu

}  // namespace zkp

struct zkp_buf {
    zkp::fr_t* d = nullptr;
    size_t n = 0;
    bool owned = true;  // false: wraps memory owned by the caller (zkp_buf_wrap)
};

struct zkp_srs {
    zkp::g1_affine* d = nullptr;  // the SRS powers P_i, i < n (96-byte points: download / trim)
    zkp::g1_tab* tab = nullptr;   // window table: row w holds 2^(c w) * P_i in 128-byte entries
    size_t n = 0;
    unsigned c = 0, W = 0;        // window bits / number of table rows (fixed at load time)
};

struct zkp_comm;   // comm.cuh

struct zkp_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    std::string last_error;
    uint64_t launches = 0;
    uint64_t msm_points = 0;  // points summed by msm_run since creation (roofline accounting)
    unsigned msm_window = 0;
    std::map<unsigned, zkp::NttDomain*> domains;
    std::map<unsigned, zkp::TwTab*> twtabs;       // four-step twiddle tables per (k, direction)
    std::map<unsigned, zkp::Coset8Tab*> coset8;   // key 8 k + u: scaling tables of coset u of the 8n domain
    zkp::fr_t* ntt_scratch = nullptr;
    size_t ntt_scratch_n = 0;
    zkp::MsmScratch* msm = nullptr;
    zkp::fr_t* io_scratch = nullptr;      // staging for the host-buffer entry points (grow-only)
    size_t io_scratch_n = 0;
    zkp::fr_t* prover_scratch = nullptr;  // prover.cu: scan / evaluation temporaries
    size_t prover_scratch_n = 0;
    void* pinned = nullptr;       // small pinned staging area for results
    size_t pinned_bytes = 0;
    // optional per-kernel timing (CUDA events on `stream`), read by bench.py for the roofline
    bool prof_on = false;
    struct ProfSpan { std::string name; cudaEvent_t a, b; };
    std::vector<ProfSpan> prof_spans;
    std::vector<cudaEvent_t> prof_pool;
    // One-shot hook the MSM calls right after it has launched its accumulate kernel: the round driver
    // uses it to queue transcript-independent work on a second stream behind an event recorded at that
    // point, so that it runs under the latency-bound bucket reduction instead of competing with the
    // accumulation for the SMs.
    void (*after_accumulate)(void*) = nullptr;
    void* after_accumulate_arg = nullptr;
};

namespace zkp {

inline int cuda_fail(zkp_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    if (ctx) ctx->last_error = buf;
    return ZKP_ERR_CUDA;
}

#define ZKP_CUDA(ctx, expr)                                                      \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) return zkp::cuda_fail(ctx, _e, #expr, __FILE__, __LINE__); \
    } while (0)

// Launch check: counts the launch and surfaces configuration errors immediately.
#define ZKP_LAUNCHED(ctx)                                                        \
    do {                                                                         \
        (ctx)->launches++;                                                       \
        cudaError_t _e = cudaGetLastError();                                     \
        if (_e != cudaSuccess) return zkp::cuda_fail(ctx, _e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// RAII span: records an event pair around the launches issued in its scope when profiling
// is enabled (zkp_prof_enable); otherwise free.
struct ProfScope {
    zkp_ctx* ctx; int idx = -1;
    ProfScope(zkp_ctx* c, const char* name) : ctx(c) {
        nvtxRangePushA(name);   // every kernel group is also an NVTX range (nsys / ncu --nvtx)
        if (!c->prof_on) return;
        cudaEvent_t a, b;
        auto get = [&](cudaEvent_t* e) {
            if (!c->prof_pool.empty()) { *e = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else cudaEventCreate(e);
        };
        get(&a); get(&b);
        cudaEventRecord(a, c->stream);
        c->prof_spans.push_back({name, a, b});
        idx = (int)c->prof_spans.size() - 1;
    }
    ~ProfScope() {
        if (idx >= 0) cudaEventRecord(ctx->prof_spans[idx].b, ctx->stream);
        nvtxRangePop();
    }
};

// NVTX range over a scope (the prover's rounds)
struct NvtxScope {
    explicit NvtxScope(const char* name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
};

inline int set_device(zkp_ctx* ctx) {
    ZKP_CUDA(ctx, cudaSetDevice(ctx->device));
    return ZKP_OK;
}

// ntt.cu
int ntt_run(zkp_ctx* ctx, const fr_t* in, size_t in_stride, size_t len_in, fr_t* out,
            size_t out_stride, unsigned k, bool inverse, bool coset, unsigned batch);
void ntt_free_domains(zkp_ctx* ctx);
int coset8_forward(zkp_ctx* ctx, const fr_t* in, size_t len_in, fr_t* out, unsigned k, unsigned first, unsigned count);
int coset8_inverse_local(zkp_ctx* ctx, fr_t* data, unsigned k, unsigned first, unsigned count);
int ntt_elements(zkp_ctx* ctx, unsigned k, fr_t* out);
fr_t fft_constant_host(unsigned k, int kind);
int ntt_permute(zkp_ctx* ctx, const fr_t* in, fr_t* out, size_t A, size_t B, size_t w);
int ntt_twiddle_transpose(zkp_ctx* ctx, const fr_t* in, fr_t* out, size_t rows, size_t cols, size_t a0, unsigned k,
                          bool inverse);
int ntt_scale_matrix(zkp_ctx* ctx, fr_t* data, size_t rows, size_t cols, size_t a0, const fr_t& base1,
                     const fr_t& base2, int mode);

// msm.cu
int msm_run(zkp_ctx* ctx, const zkp_srs* srs, const fr_t* scalars_dev, size_t n, g1_affine* out_host);
int msm_run_batch(zkp_ctx* ctx, const zkp_srs* srs, const fr_t* const* scalars_dev, const size_t* lens, unsigned nb,
                  g1_affine* out_host, int* overflow);
int msm_run_batch_ex(zkp_ctx* ctx, zkp_comm* cm, const zkp_srs* srs, const fr_t* const* scalars_dev, const size_t* lens,
                     const size_t* offs, unsigned nb, g1_affine* out_host, int* overflow);
int msm_commit_sharded(zkp_ctx* ctx, zkp_comm* cm, const zkp_srs* srs, const fr_t* const* polys, const size_t* lens,
                       unsigned nb, g1_affine* out_host, int* overflow);
unsigned msm_choose_window(size_t n);
int srs_build_table(zkp_ctx* ctx, zkp_srs* srs);
int msm_highest_nonzero(zkp_ctx* ctx, const fr_t* scalars_dev, size_t n, long long* out);
int srs_generate(zkp_ctx* ctx, const fr_t& tau, size_t first, size_t n, g1_affine* out_dev);
void msm_free(zkp_ctx* ctx);

// prover.cu
void prover_free(zkp_ctx* ctx);
int poly_eval2_launch(zkp_ctx* ctx, const zkp_poly_ref* polys, const uint8_t* which, unsigned count,
                      const uint64_t points[8], fr_t** res_dev);

}  // namespace zkp

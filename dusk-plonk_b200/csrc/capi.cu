// extern "C" boundary of libzkp_b200.so (declared in include/zkp_b200.h).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "comm.cuh"

using namespace zkp;

static_assert(sizeof(fr_t) == 32, "Fr is 4 x u64");
static_assert(sizeof(fq_t) == 48, "Fq is 6 x u64");
static_assert(sizeof(g1_affine) == 96, "G1 affine is 12 x u64");

extern "C" {

const char* zkp_strerror(int code) {
    switch (code) {
        case ZKP_OK: return "ok";
        case ZKP_ERR_INVALID: return "invalid argument";
        case ZKP_ERR_CUDA: return "CUDA failure";
        case ZKP_ERR_DEGREE: return "polynomial degree exceeds the SRS";
        case ZKP_ERR_NOMEM: return "out of memory";
        case ZKP_ERR_VERIFY: return "proof rejected";
        case ZKP_ERR_STATE: return "call out of order (prove_witness before set_wiring)";
        default: return "unknown error";
    }
}

int zkp_ctx_create(int device, zkp_ctx** out) {
    if (!out) return ZKP_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return ZKP_ERR_CUDA;
    zkp_ctx* ctx = new zkp_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        prop.major < 10) {  // kernels are built for sm_100a only; no fallback
        delete ctx;
        return ZKP_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev_start) != cudaSuccess || cudaEventCreate(&ctx->ev_stop) != cudaSuccess) {
        delete ctx;
        return ZKP_ERR_CUDA;
    }
    ctx->pinned_bytes = 1 << 16;
    if (cudaMallocHost(&ctx->pinned, ctx->pinned_bytes) != cudaSuccess) {
        delete ctx;
        return ZKP_ERR_CUDA;
    }
    *out = ctx;
    return ZKP_OK;
}

void zkp_ctx_destroy(zkp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ntt_free_domains(ctx);
    if (ctx->io_scratch) cudaFree(ctx->io_scratch);
    msm_free(ctx);
    prover_free(ctx);
    for (auto& s : ctx->prof_spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto& e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaEventDestroy(ctx->ev_start);
    cudaEventDestroy(ctx->ev_stop);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* zkp_last_error(const zkp_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int zkp_ctx_sync(zkp_ctx* ctx) {
    if (!ctx) return ZKP_ERR_INVALID;
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

void* zkp_ctx_stream(zkp_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int zkp_sm_count(const zkp_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t zkp_launch_count(const zkp_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t zkp_msm_point_count(const zkp_ctx* ctx) { return ctx ? ctx->msm_points : 0; }

int zkp_timer_start(zkp_ctx* ctx) {
    if (!ctx) return ZKP_ERR_INVALID;
    ZKP_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return ZKP_OK;
}

int zkp_timer_stop_ms(zkp_ctx* ctx, float* ms) {
    if (!ctx || !ms) return ZKP_ERR_INVALID;
    ZKP_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    ZKP_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
    ZKP_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
    return ZKP_OK;
}

int zkp_prof_enable(zkp_ctx* ctx, int on) {
    if (!ctx) return ZKP_ERR_INVALID;
    ctx->prof_on = on != 0;
    return ZKP_OK;
}

int zkp_prof_reset(zkp_ctx* ctx) {
    if (!ctx) return ZKP_ERR_INVALID;
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& s : ctx->prof_spans) { ctx->prof_pool.push_back(s.a); ctx->prof_pool.push_back(s.b); }
    ctx->prof_spans.clear();
    return ZKP_OK;
}

int zkp_prof_read(zkp_ctx* ctx, const char* name, float* total_ms, uint64_t* count) {
    if (!ctx || !name || !total_ms || !count) return ZKP_ERR_INVALID;
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float tot = 0;
    uint64_t n = 0;
    for (auto& s : ctx->prof_spans) {
        if (s.name != name) continue;
        float ms = 0;
        ZKP_CUDA(ctx, cudaEventElapsedTime(&ms, s.a, s.b));
        tot += ms;
        n++;
    }
    *total_ms = tot;
    *count = n;
    return ZKP_OK;
}

/* ---- buffers ---------------------------------------------------------------------- */
int zkp_buf_alloc(zkp_ctx* ctx, size_t n, zkp_buf** out) {
    if (!ctx || !out) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    zkp_buf* b = new zkp_buf();
    b->n = n;
    cudaError_t e = cudaMalloc(&b->d, (n ? n : 1) * sizeof(fr_t));
    if (e != cudaSuccess) {
        delete b;
        cuda_fail(ctx, e, "cudaMalloc", __FILE__, __LINE__);
        return e == cudaErrorMemoryAllocation ? ZKP_ERR_NOMEM : ZKP_ERR_CUDA;
    }
    *out = b;
    return ZKP_OK;
}

int zkp_buf_free(zkp_ctx* ctx, zkp_buf* buf) {
    if (!ctx || !buf) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (buf->owned) ZKP_CUDA(ctx, cudaFree(buf->d));
    delete buf;
    return ZKP_OK;
}

int zkp_buf_wrap(zkp_ctx* ctx, void* device_ptr, size_t n, zkp_buf** out) {
    if (!ctx || !out || (!device_ptr && n) || ((uintptr_t)device_ptr & 15)) return ZKP_ERR_INVALID;
    zkp_buf* b = new zkp_buf();
    b->d = reinterpret_cast<fr_t*>(device_ptr);
    b->n = n;
    b->owned = false;
    *out = b;
    return ZKP_OK;
}

int zkp_permute_dev(zkp_ctx* ctx, const zkp_buf* in, size_t in_off, zkp_buf* out, size_t out_off, size_t A, size_t B,
                    size_t w) {
    if (!ctx || !in || !out) return ZKP_ERR_INVALID;
    const size_t total = A * B * w;
    if (in_off + total > in->n || out_off + total > out->n || in->d + in_off == out->d + out_off) return ZKP_ERR_INVALID;
    return ntt_permute(ctx, in->d + in_off, out->d + out_off, A, B, w);
}

int zkp_scale_matrix_dev(zkp_ctx* ctx, zkp_buf* data, size_t off, size_t rows, size_t cols, size_t a0,
                         const uint64_t base1[4], const uint64_t base2[4], int mode) {
    if (!ctx || !data || !base1 || !base2 || off + rows * cols > data->n || mode < 0 || mode > 1) return ZKP_ERR_INVALID;
    fr_t b1, b2;
    memcpy(b1.l, base1, 32);
    memcpy(b2.l, base2, 32);
    return ntt_scale_matrix(ctx, data->d + off, rows, cols, a0, b1, b2, mode);
}

int zkp_twiddle_transpose_dev(zkp_ctx* ctx, const zkp_buf* in, size_t in_off, zkp_buf* out, size_t out_off, size_t rows,
                              size_t cols, size_t a0, unsigned k, int inverse) {
    if (!ctx || !in || !out) return ZKP_ERR_INVALID;
    const size_t total = rows * cols;
    if (in_off + total > in->n || out_off + total > out->n || in->d + in_off == out->d + out_off) return ZKP_ERR_INVALID;
    return ntt_twiddle_transpose(ctx, in->d + in_off, out->d + out_off, rows, cols, a0, k, inverse != 0);
}

/* strided host <-> device copies: `height` rows of `width` Fr; the host side has a row pitch of `host_pitch` Fr,
 * the device side is dense.  (A rank's column slab of a natural-order host vector) */
int zkp_buf_upload_2d(zkp_ctx* ctx, zkp_buf* dst, size_t dst_off, const uint64_t* src, size_t width, size_t height,
                      size_t host_pitch) {
    if (!ctx || !dst || (!src && width * height) || dst_off + width * height > dst->n || host_pitch < width)
        return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (width * height == 0) return ZKP_OK;
    ZKP_CUDA(ctx, cudaMemcpy2DAsync(dst->d + dst_off, width * sizeof(fr_t), src, host_pitch * sizeof(fr_t),
                                    width * sizeof(fr_t), height, cudaMemcpyHostToDevice, ctx->stream));
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

int zkp_buf_download_2d(zkp_ctx* ctx, const zkp_buf* src, size_t src_off, uint64_t* dst, size_t width, size_t height,
                        size_t host_pitch) {
    if (!ctx || !src || (!dst && width * height) || src_off + width * height > src->n || host_pitch < width)
        return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    if (width * height == 0) return ZKP_OK;
    ZKP_CUDA(ctx, cudaMemcpy2DAsync(dst, host_pitch * sizeof(fr_t), src->d + src_off, width * sizeof(fr_t),
                                    width * sizeof(fr_t), height, cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

size_t zkp_buf_len(const zkp_buf* buf) { return buf ? buf->n : 0; }

int zkp_buf_upload(zkp_ctx* ctx, zkp_buf* dst, size_t dst_off, const uint64_t* src, size_t n) {
    if (!ctx || !dst || (!src && n) || dst_off + n > dst->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(dst->d + dst_off, src, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

int zkp_buf_download(zkp_ctx* ctx, const zkp_buf* src, size_t src_off, uint64_t* dst, size_t n) {
    if (!ctx || !src || (!dst && n) || src_off + n > src->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(dst, src->d + src_off, n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

int zkp_buf_zero(zkp_ctx* ctx, zkp_buf* buf, size_t off, size_t n) {
    if (!ctx || !buf || off + n > buf->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    ZKP_CUDA(ctx, cudaMemsetAsync(buf->d + off, 0, n * sizeof(fr_t), ctx->stream));
    return ZKP_OK;
}

int zkp_buf_copy(zkp_ctx* ctx, zkp_buf* dst, size_t dst_off, const zkp_buf* src, size_t src_off, size_t n) {
    if (!ctx || !dst || !src || dst_off + n > dst->n || src_off + n > src->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(dst->d + dst_off, src->d + src_off, n * sizeof(fr_t), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    return ZKP_OK;
}

/* Keccak-f[1600] for the host-side Merlin/STROBE transcript (TranscriptProtocol stays on the
 * host; this is plain C so the Python mirror does not spend its time in a bytecode loop). */
void zkp_keccak_f1600(uint64_t st[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
        0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
        0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39,
                                41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int r = 0; r < 24; r++) {
        uint64_t c[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int x = 0; x < 5; x++) {
            const uint64_t t = c[(x + 1) % 5];
            const uint64_t d = c[(x + 4) % 5] ^ ((t << 1) | (t >> 63));
            for (int y = 0; y < 25; y += 5) st[x + y] ^= d;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) {
                const uint64_t v = st[x + 5 * y];
                const int k = ROT[x + 5 * y];
                b[y + 5 * ((2 * x + 3 * y) % 5)] = k ? ((v << k) | (v >> (64 - k))) : v;
            }
        for (int y = 0; y < 25; y += 5)
            for (int x = 0; x < 5; x++) st[x + y] = b[x + y] ^ (~b[(x + 1) % 5 + y] & b[(x + 2) % 5 + y]);
        st[0] ^= RC[r];
    }
}

int zkp_host_alloc(size_t bytes, void** out) {
    if (!out) return ZKP_ERR_INVALID;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { *out = nullptr; return e == cudaErrorMemoryAllocation ? ZKP_ERR_NOMEM : ZKP_ERR_CUDA; }
    return ZKP_OK;
}

int zkp_host_free(void* p) {
    if (!p) return ZKP_OK;
    return cudaFreeHost(p) == cudaSuccess ? ZKP_OK : ZKP_ERR_CUDA;
}

/* ---- NTT -------------------------------------------------------------------------- */
int zkp_ntt_dev_batch(zkp_ctx* ctx, const zkp_buf* in, size_t in_stride, size_t len_in, zkp_buf* out,
                      size_t out_stride, unsigned k, int inverse, int coset, unsigned batch) {
    if (!ctx || !in || !out || k > 28 || batch == 0) return ZKP_ERR_INVALID;
    const size_t n = (size_t)1 << k;
    if (len_in > n) return ZKP_ERR_INVALID;
    if ((batch - 1) * in_stride + len_in > in->n || (batch - 1) * out_stride + n > out->n) return ZKP_ERR_INVALID;
    if (batch > 1 && (in_stride < len_in || out_stride < n)) return ZKP_ERR_INVALID;
    return ntt_run(ctx, in->d, in_stride, len_in, out->d, out_stride, k, inverse != 0, coset != 0, batch);
}

int zkp_ntt_dev(zkp_ctx* ctx, const zkp_buf* in, size_t len_in, zkp_buf* out, unsigned k, int inverse,
                int coset) {
    return zkp_ntt_dev_batch(ctx, in, 0, len_in, out, 0, k, inverse, coset, 1);
}

// Device staging area of the host-buffer entry points: grown on demand, kept for the life of the
// context so that repeated calls do not pay cudaMalloc / cudaFree (both synchronise the device).
static int io_buffer(zkp_ctx* ctx, size_t n, fr_t** out) {
    if (ctx->io_scratch_n < n) {
        if (ctx->io_scratch) cudaFree(ctx->io_scratch);
        ctx->io_scratch = nullptr;
        ctx->io_scratch_n = 0;
        cudaError_t e = cudaMalloc(&ctx->io_scratch, n * sizeof(fr_t));
        if (e != cudaSuccess) {
            cuda_fail(ctx, e, "cudaMalloc(io staging)", __FILE__, __LINE__);
            return e == cudaErrorMemoryAllocation ? ZKP_ERR_NOMEM : ZKP_ERR_CUDA;
        }
        ctx->io_scratch_n = n;
    }
    *out = ctx->io_scratch;
    return ZKP_OK;
}

int zkp_ntt(zkp_ctx* ctx, uint64_t* data, size_t len_in, unsigned k, int inverse, int coset) {
    if (!ctx || !data || k > 28) return ZKP_ERR_INVALID;
    const size_t n = (size_t)1 << k;
    if (len_in > n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    fr_t* buf = nullptr;
    if ((rc = io_buffer(ctx, n, &buf))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(buf, data, len_in * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = ntt_run(ctx, buf, 0, len_in, buf, 0, k, inverse != 0, coset != 0, 1))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(data, buf, n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

int zkp_fft_constant(unsigned k, int kind, uint64_t out[4]) {
    if (k > 32 || kind < 0 || kind > 4 || !out) return ZKP_ERR_INVALID;
    fr_t v = fft_constant_host(k, kind);
    memcpy(out, v.l, 32);
    return ZKP_OK;
}

int zkp_fft_elements_dev(zkp_ctx* ctx, unsigned k, zkp_buf* out) {
    if (!ctx || !out || k > 28 || out->n < ((size_t)1 << k)) return ZKP_ERR_INVALID;
    return ntt_elements(ctx, k, out->d);
}

/* ---- SRS / MSM -------------------------------------------------------------------- */
// Allocates the W-row table for n powers; row 0 is filled by the caller.
static int srs_alloc(zkp_ctx* ctx, size_t n, zkp_srs** out) {
    zkp_srs* s = new zkp_srs();
    s->n = n;
    unsigned forced = ctx->msm_window;
    if (!forced)
        if (const char* e = getenv("ZKP_MSM_WINDOW")) forced = (unsigned)atoi(e);  // tuning knob
    if (forced > 22) forced = 22;
    s->c = forced ? (forced < 2 ? 2 : forced) : msm_choose_window(n);
    s->W = 255 / s->c + 1;
    if ((size_t)s->W * n >= (1ull << 31)) { delete s; return ZKP_ERR_INVALID; }
    cudaError_t e = cudaMalloc(&s->d, (n ? n : 1) * sizeof(g1_affine));
    if (e != cudaSuccess) {
        delete s;
        cuda_fail(ctx, e, "cudaMalloc(srs table)", __FILE__, __LINE__);
        return e == cudaErrorMemoryAllocation ? ZKP_ERR_NOMEM : ZKP_ERR_CUDA;
    }
    *out = s;
    return ZKP_OK;
}

int zkp_srs_load(zkp_ctx* ctx, const uint64_t* xy, size_t n, zkp_srs** out) {
    if (!ctx || !out || (!xy && n)) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    zkp_srs* s = nullptr;
    if ((rc = srs_alloc(ctx, n, &s))) return rc;
    cudaError_t e = cudaMemcpyAsync(s->d, xy, n * sizeof(g1_affine), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(s->d);
        delete s;
        return cuda_fail(ctx, e, "srs upload", __FILE__, __LINE__);
    }
    if ((rc = srs_build_table(ctx, s))) { cudaFree(s->d); cudaFree(s->tab); delete s; return rc; }
    *out = s;
    return ZKP_OK;
}

int zkp_srs_generate(zkp_ctx* ctx, const uint64_t tau[4], size_t n, zkp_srs** out) {
    return zkp_srs_generate_range(ctx, tau, 0, n, out);
}

int zkp_srs_generate_range(zkp_ctx* ctx, const uint64_t tau[4], size_t first, size_t n, zkp_srs** out) {
    if (!ctx || !out || !tau) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    zkp_srs* s = nullptr;
    if ((rc = srs_alloc(ctx, n, &s))) return rc;
    fr_t t;
    memcpy(t.l, tau, 32);
    rc = srs_generate(ctx, t, first, n, s->d);
    if (!rc) rc = srs_build_table(ctx, s);
    if (rc) { cudaFree(s->d); cudaFree(s->tab); delete s; return rc; }
    *out = s;
    return ZKP_OK;
}

/* PlonkParams::trim (src/key.rs:82) without leaving the device: the first `keep` powers of an SRS as a new one.
 * When the window width stays the same the table rows are sliced (device-to-device copies), else rebuilt. */
int zkp_srs_trim(zkp_ctx* ctx, const zkp_srs* srs, size_t keep, zkp_srs** out) {
    if (!ctx || !srs || !out || keep > srs->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    zkp_srs* s = nullptr;
    if ((rc = srs_alloc(ctx, keep, &s))) return rc;
    cudaError_t e = cudaMemcpyAsync(s->d, srs->d, keep * sizeof(g1_affine), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess && keep && s->c == srs->c && srs->tab) {
        e = cudaMalloc(&s->tab, (size_t)s->W * keep * sizeof(g1_tab));
        for (unsigned w = 0; e == cudaSuccess && w < s->W; w++)
            e = cudaMemcpyAsync(s->tab + (size_t)w * keep, srs->tab + (size_t)w * srs->n, keep * sizeof(g1_tab),
                                cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    } else if (e == cudaSuccess) {
        rc = srs_build_table(ctx, s);
    }
    if (e != cudaSuccess || rc) {
        cudaFree(s->d); cudaFree(s->tab); delete s;
        return rc ? rc : cuda_fail(ctx, e, "srs trim", __FILE__, __LINE__);
    }
    *out = s;
    return ZKP_OK;
}

int zkp_srs_download(zkp_ctx* ctx, const zkp_srs* srs, size_t off, uint64_t* xy, size_t n) {
    if (!ctx || !srs || (!xy && n) || off + n > srs->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(xy, srs->d + off, n * sizeof(g1_affine), cudaMemcpyDeviceToHost, ctx->stream));
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZKP_OK;
}

int zkp_srs_free(zkp_ctx* ctx, zkp_srs* srs) {
    if (!ctx || !srs) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    ZKP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZKP_CUDA(ctx, cudaFree(srs->d));
    if (srs->tab) ZKP_CUDA(ctx, cudaFree(srs->tab));
    delete srs;
    return ZKP_OK;
}

size_t zkp_srs_len(const zkp_srs* srs) { return srs ? srs->n : 0; }

int zkp_msm_set_window(zkp_ctx* ctx, unsigned c) {
    if (!ctx || c > 22) return ZKP_ERR_INVALID;
    ctx->msm_window = c;
    return ZKP_OK;
}

int zkp_msm_g1_dev(zkp_ctx* ctx, const zkp_srs* srs, const zkp_buf* scalars, size_t off, size_t n,
                   uint64_t out_xy[12]) {
    if (!ctx || !srs || !scalars || !out_xy || off + n > scalars->n || n > srs->n) return ZKP_ERR_INVALID;
    return msm_run(ctx, srs, scalars->d + off, n, reinterpret_cast<g1_affine*>(out_xy));
}

int zkp_commit_dev(zkp_ctx* ctx, const zkp_srs* srs, const zkp_buf* coeffs, size_t off, size_t n,
                   uint64_t out_xy[12]) {
    if (!ctx || !srs || !coeffs || !out_xy || off + n > coeffs->n) return ZKP_ERR_INVALID;
    const fr_t* p = coeffs->d + off;
    int ovf = 0;
    int rc = msm_run_batch(ctx, srs, &p, &n, 1, reinterpret_cast<g1_affine*>(out_xy), &ovf);
    if (rc) return rc;
    return ovf ? ZKP_ERR_DEGREE : ZKP_OK;
}

int zkp_commit_batch_dev(zkp_ctx* ctx, const zkp_srs* srs, const zkp_poly_ref* polys, unsigned count,
                         uint64_t* out_xy, int* status) {
    if (!ctx || !srs || !polys || !out_xy || !status || count == 0 || count > 8) return ZKP_ERR_INVALID;
    const fr_t* ptrs[8];
    size_t lens[8];
    int ovf[8];
    for (unsigned i = 0; i < count; i++) {
        if (!polys[i].buf || polys[i].off + polys[i].len > polys[i].buf->n) return ZKP_ERR_INVALID;
        ptrs[i] = polys[i].buf->d + polys[i].off;
        lens[i] = polys[i].len;
    }
    int rc = msm_run_batch(ctx, srs, ptrs, lens, count, reinterpret_cast<g1_affine*>(out_xy), ovf);
    if (rc) return rc;
    for (unsigned i = 0; i < count; i++) status[i] = ovf[i] ? ZKP_ERR_DEGREE : ZKP_OK;
    return ZKP_OK;
}

int zkp_commit_batch_sharded_dev(zkp_ctx* ctx, zkp_comm* comm, const zkp_srs* srs, const zkp_poly_ref* polys,
                                 unsigned count, uint64_t* out_xy, int* status) {
    if (!ctx || !srs || !polys || !out_xy || !status || count == 0 || count > 8) return ZKP_ERR_INVALID;
    const fr_t* ptrs[8];
    size_t lens[8];
    int ovf[8];
    for (unsigned i = 0; i < count; i++) {
        if (!polys[i].buf || polys[i].off + polys[i].len > polys[i].buf->n) return ZKP_ERR_INVALID;
        ptrs[i] = polys[i].buf->d + polys[i].off;
        lens[i] = polys[i].len;
    }
    int rc = msm_commit_sharded(ctx, comm, srs, ptrs, lens, count, reinterpret_cast<g1_affine*>(out_xy), ovf);
    if (rc) return rc;
    for (unsigned i = 0; i < count; i++) status[i] = ovf[i] ? ZKP_ERR_DEGREE : ZKP_OK;
    return ZKP_OK;
}

int zkp_coset8_ntt_dev(zkp_ctx* ctx, const zkp_buf* in, size_t in_off, size_t len_in, zkp_buf* out, size_t out_off,
                       unsigned k, unsigned first, unsigned count) {
    if (!ctx || !in || !out || k > 25 || count == 0 || first + count > 8) return ZKP_ERR_INVALID;
    const size_t n = (size_t)1 << k;
    if (in_off + len_in > in->n || len_in > 2 * n || out_off + (size_t)count * n > out->n) return ZKP_ERR_INVALID;
    return coset8_forward(ctx, in->d + in_off, len_in, out->d + out_off, k, first, count);
}

int zkp_poly_degree_dev(zkp_ctx* ctx, const zkp_buf* coeffs, size_t off, size_t n, long long* top) {
    if (!ctx || !coeffs || !top || off + n > coeffs->n) return ZKP_ERR_INVALID;
    return msm_highest_nonzero(ctx, coeffs->d + off, n, top);
}

int zkp_msm_g1(zkp_ctx* ctx, const zkp_srs* srs, const uint64_t* scalars, size_t n, uint64_t out_xy[12]) {
    if (!ctx || !srs || (!scalars && n) || !out_xy || n > srs->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    fr_t* buf = nullptr;
    if ((rc = io_buffer(ctx, n ? n : 1, &buf))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(buf, scalars, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    return msm_run(ctx, srs, buf, n, reinterpret_cast<g1_affine*>(out_xy));
}

int zkp_commit(zkp_ctx* ctx, const zkp_srs* srs, const uint64_t* coeffs, size_t n, uint64_t out_xy[12]) {
    if (!ctx || !srs || (!coeffs && n) || !out_xy) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    fr_t* buf = nullptr;
    if ((rc = io_buffer(ctx, n ? n : 1, &buf))) return rc;
    ZKP_CUDA(ctx, cudaMemcpyAsync(buf, coeffs, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    const fr_t* p = buf;
    int ovf = 0;
    rc = msm_run_batch(ctx, srs, &p, &n, 1, reinterpret_cast<g1_affine*>(out_xy), &ovf);
    if (rc) return rc;
    return ovf ? ZKP_ERR_DEGREE : ZKP_OK;
}

}  // extern "C"

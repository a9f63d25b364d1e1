// Multi-GPU plumbing behind the C ABI: one process (or thread) per GPU, NCCL over NVLink / NVSwitch.
//
// The reference has no multi-GPU path; SURVEY 8(e) / BASELINE north_star define the split: commitments by
// SRS ranges with a partial-sum gather, the 8n-point quotient domain by cosets (one exchange for the inverse
// transform).  A Rust host binds zkp_comm_* exactly like the rest of the ABI: rank 0 calls
// zkp_comm_unique_id, ships the 256 bytes to its peers by whatever channel it has, every rank calls
// zkp_comm_create.  NCCL is resolved at run time (dlopen) so that the library loads -- and every
// single-GPU entry point works -- on a box without it, and so that a host process that already carries its
// own NCCL (PyTorch bundles one) shares that copy instead of loading a second one.
#include <dlfcn.h>
#include <nccl.h>      // types and prototypes only; nothing links against libnccl
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "comm.cuh"

namespace zkp {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static bool nccl_load(std::string* err) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return true;
    const char* names[3] = {getenv("ZKP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { *err = std::string("cannot load NCCL: ") + (dlerror() ? dlerror() : "not found"); return false; }
#define ZKP_SYM(field, name)                                                              \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));              \
    if (!g_nccl.field) { *err = std::string("NCCL symbol missing: ") + name; dlclose(h); return false; }
    ZKP_SYM(GetUniqueId, "ncclGetUniqueId")
    ZKP_SYM(CommInitRank, "ncclCommInitRank")
    ZKP_SYM(CommDestroy, "ncclCommDestroy")
    ZKP_SYM(AllGather, "ncclAllGather")
    ZKP_SYM(Send, "ncclSend")
    ZKP_SYM(Recv, "ncclRecv")
    ZKP_SYM(GroupStart, "ncclGroupStart")
    ZKP_SYM(GroupEnd, "ncclGroupEnd")
    ZKP_SYM(GetErrorString, "ncclGetErrorString")
    ZKP_SYM(GetVersion, "ncclGetVersion")
#undef ZKP_SYM
    g_nccl.handle = h;
    return true;
}

static int nccl_fail(zkp_ctx* ctx, ncclResult_t r, const char* what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    if (ctx) ctx->last_error = buf;
    return ZKP_ERR_CUDA;
}
#define ZKP_NCCL(ctx, expr)                                              \
    do {                                                                 \
        ncclResult_t _r = (expr);                                        \
        if (_r != ncclSuccess) return nccl_fail(ctx, _r, #expr);         \
    } while (0)

// every rank contributes `bytes` from send (device) and receives nranks * bytes into recv (device)
int comm_allgather(zkp_comm* cm, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    if (!cm || cm->nranks == 1) {
        if (send != recv) ZKP_CUDA(cm ? cm->ctx : nullptr, cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st));
        return ZKP_OK;
    }
    ZKP_NCCL(cm->ctx, g_nccl.AllGather(send, recv, bytes, ncclUint8, static_cast<ncclComm_t>(cm->nccl), st));
    cm->collectives++;
    cm->bytes_sent += bytes * (size_t)(cm->nranks - 1);
    return ZKP_OK;
}

// Coset exchange of the distributed inverse transform: this rank holds `local` vectors of `n` elements
// (its cosets, in order); slab s of every one of them goes to rank s, and from every rank r this rank
// receives slab `rank` of r's vectors into recv[(r * local + j) * slab], slab = n / nranks elements.
int comm_exchange_slabs(zkp_comm* cm, const fr_t* send, fr_t* recv, size_t n, unsigned local, cudaStream_t st) {
    const int G = cm ? cm->nranks : 1;
    const size_t slab = n / (size_t)G;
    if (G == 1) {
        ZKP_CUDA(cm ? cm->ctx : nullptr, cudaMemcpyAsync(recv, send, (size_t)local * n * sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
        return ZKP_OK;
    }
    ncclComm_t nc = static_cast<ncclComm_t>(cm->nccl);
    ZKP_NCCL(cm->ctx, g_nccl.GroupStart());
    for (int peer = 0; peer < G; peer++) {
        for (unsigned j = 0; j < local; j++) {
            ZKP_NCCL(cm->ctx, g_nccl.Send(send + (size_t)j * n + (size_t)peer * slab, slab * sizeof(fr_t), ncclUint8, peer, nc, st));
            ZKP_NCCL(cm->ctx, g_nccl.Recv(recv + ((size_t)peer * local + j) * slab, slab * sizeof(fr_t), ncclUint8, peer, nc, st));
        }
    }
    ZKP_NCCL(cm->ctx, g_nccl.GroupEnd());
    cm->collectives++;
    cm->bytes_sent += (size_t)local * (size_t)(G - 1) * slab * sizeof(fr_t);
    return ZKP_OK;
}

}  // namespace zkp

using namespace zkp;

extern "C" {

int zkp_comm_unique_id(uint8_t out[256]) {
    if (!out) return ZKP_ERR_INVALID;
    std::string err;
    if (!nccl_load(&err)) return ZKP_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) <= 128, "unique id fits its slot");
    memset(out, 0, 256);
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ZKP_ERR_CUDA;
    memcpy(out, &id, sizeof id);      // bytes 128 .. 255 are reserved (zero)
    return ZKP_OK;
}

int zkp_comm_create(zkp_ctx* ctx, const uint8_t id[256], int rank, int nranks, zkp_comm** out) {
    if (!ctx || !out || rank < 0 || rank >= nranks) return ZKP_ERR_INVALID;
    if (nranks != 1 && nranks != 2 && nranks != 4 && nranks != 8) return ZKP_ERR_INVALID;   // cosets of the 8n domain
    if (nranks > 1 && !id) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(ctx))) return rc;
    zkp_comm* cm = new zkp_comm();
    cm->ctx = ctx; cm->rank = rank; cm->nranks = nranks;
    if (nranks > 1) {
        std::string err;
        if (!nccl_load(&err)) { ctx->last_error = err; delete cm; return ZKP_ERR_CUDA; }
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof uid);
        ncclComm_t nc = nullptr;
        ncclResult_t r = g_nccl.CommInitRank(&nc, nranks, uid, rank);
        if (r != ncclSuccess) { delete cm; return nccl_fail(ctx, r, "ncclCommInitRank"); }
        cm->nccl = nc;
    }
    // gather staging: per rank 8 partial sums (XYZZ) + flags / scalars
    cm->slot_bytes = 8 * sizeof(g1_xyzz) + 512;
    if (cudaMalloc(&cm->gsend, cm->slot_bytes) != cudaSuccess ||
        cudaMalloc(&cm->grecv, cm->slot_bytes * (size_t)nranks) != cudaSuccess ||
        cudaMallocHost(&cm->hrecv, cm->slot_bytes * (size_t)nranks) != cudaSuccess) {
        zkp_comm_destroy(cm);
        return ZKP_ERR_NOMEM;
    }
    *out = cm;
    return ZKP_OK;
}

int zkp_comm_destroy(zkp_comm* cm) {
    if (!cm) return ZKP_OK;
    if (cm->ctx) { cudaSetDevice(cm->ctx->device); cudaStreamSynchronize(cm->ctx->stream); }
    if (cm->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(cm->nccl));
    if (cm->gsend) cudaFree(cm->gsend);
    if (cm->grecv) cudaFree(cm->grecv);
    if (cm->hrecv) cudaFreeHost(cm->hrecv);
    delete cm;
    return ZKP_OK;
}

int zkp_comm_rank(const zkp_comm* cm) { return cm ? cm->rank : 0; }
int zkp_comm_size(const zkp_comm* cm) { return cm ? cm->nranks : 1; }

/* all-to-all of equal blocks between device vectors: block p of src goes to rank p, block r of dst comes from
 * rank r (count Fr elements each); on the context's stream, no host synchronisation */
int zkp_comm_all_to_all_dev(zkp_comm* cm, const zkp_buf* src, size_t src_off, zkp_buf* dst, size_t dst_off, size_t count) {
    if (!cm || !src || !dst) return ZKP_ERR_INVALID;
    const size_t G = (size_t)cm->nranks;
    if (src_off + G * count > src->n || dst_off + G * count > dst->n) return ZKP_ERR_INVALID;
    int rc;
    if ((rc = set_device(cm->ctx))) return rc;
    // one "coset" of G * count elements whose slab p goes to rank p: exactly comm_exchange_slabs with local = 1
    return comm_exchange_slabs(cm, src->d + src_off, dst->d + dst_off, G * count, 1, cm->ctx->stream);
}

/* collectives issued / bytes this rank put on the wire since creation (bench.py reports them) */
int zkp_comm_stats(const zkp_comm* cm, uint64_t* collectives, uint64_t* bytes_sent) {
    if (!cm || !collectives || !bytes_sent) return ZKP_ERR_INVALID;
    *collectives = cm->collectives;
    *bytes_sent = cm->bytes_sent;
    return ZKP_OK;
}

}  // extern "C"

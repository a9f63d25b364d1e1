// zkp_comm: one rank's view of the GPUs that share a job (comm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

struct zkp_comm {
    zkp_ctx* ctx = nullptr;
    int rank = 0, nranks = 1;
    void* nccl = nullptr;          // ncclComm_t (nullptr when nranks == 1)
    // staging of the small gathers (partial commitments, flags, partial evaluations)
    size_t slot_bytes = 0;
    uint8_t* gsend = nullptr;      // device, slot_bytes
    uint8_t* grecv = nullptr;      // device, nranks * slot_bytes
    uint8_t* hrecv = nullptr;      // pinned host mirror of grecv
    uint64_t collectives = 0, bytes_sent = 0;
};

namespace zkp {
int comm_allgather(zkp_comm* cm, const void* send, void* recv, size_t bytes, cudaStream_t st);
int comm_exchange_slabs(zkp_comm* cm, const fr_t* send, fr_t* recv, size_t n, unsigned local, cudaStream_t st);
}  // namespace zkp

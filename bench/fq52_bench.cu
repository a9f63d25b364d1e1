// Throughput of the FP64-pipe Fq multiplication (bench/fq52.cuh) on sm_100a, alone and side by side with
// the product's IMAD multiplication (arith.cuh) in alternating warps.  EXPERIMENT for round 2.
//   mode "imad":   every warp runs the 32-bit-limb Montgomery multiplication   (the round-1 baseline)
//   mode "dfma":   every warp runs the 52-bit-limb FP64 multiplication
//   mode "hybrid": warps alternate (ratio given), each kind on its own pipes; total multiplications / s
#include <cuda_runtime.h>
#include <stdio.h>

#include "../dusk-plonk_b200/csrc/arith.cuh"
#include "fq52.cuh"
#include "fq48.cuh"

using namespace zkp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// fq48: the FP64-only multiplication (no integer accumulation at all)
template <int DFMA_OF>
__global__ void __launch_bounds__(128) k_mix48(uint32_t* out, int iters) {
    const unsigned warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    if ((int)warp < DFMA_OF) {
        fq48::el a, b;
#pragma unroll
        for (int i = 0; i < 8; i++) { a.l[i] = (double)(1000 + i + threadIdx.x); b.l[i] = (double)(77 + 3 * i + blockIdx.x % 100); }
        for (int it = 0; it < iters; it++) { a = fq48::mul(a, b); b = fq48::mul(b, a); }
#pragma unroll
        for (int i = 0; i < 8; i++) acc ^= (uint32_t)(long long)(a.l[i] + b.l[i]);
    } else {
        fq_t a = fq_t::one(), b = fq_t::r2();
        a.l[0] ^= threadIdx.x; b.l[1] ^= blockIdx.x;
        for (int it = 0; it < iters; it++) { a = a * b; b = b * a; }
        const fq_t s = a + b;
#pragma unroll
        for (int i = 0; i < 12; i++) acc ^= s.l[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// device exactness of fq48 against the integer path (both: a b 2^-384 mod p)
__global__ void k_check48(int* bad) {
    fq_t a = fq_t::r2(), b = fq_t::one();
    a.l[0] ^= threadIdx.x * 2654435761u; b.l[3] ^= blockIdx.x * 40503u + threadIdx.x;
    a.l[11] &= 0x0fffffffu; b.l[11] &= 0x0fffffffu;
    auto to48 = [](const fq_t& c) {
        fq48::el r;
        for (int i = 0; i < 8; i++) {
            const int bit = 48 * i, w = bit / 32, off = bit % 32;
            unsigned long long lo = c.l[w] | ((unsigned long long)(w + 1 < 12 ? c.l[w + 1] : 0u) << 32);
            unsigned long long v = lo >> off;
            if (off > 16 && w + 2 < 12) v |= (unsigned long long)c.l[w + 2] << (64 - off);
            r.l[i] = (double)(v & 0xffffffffffffull);
        }
        return r;
    };
    for (int it = 0; it < 8; it++) {
        const fq_t ref = a * b;
        const fq48::el got = fq48::mul(to48(a), to48(b));
        const fq48::el want = to48(ref);
        for (int i = 0; i < 8; i++) if (got.l[i] != want.l[i]) { atomicAdd(bad, 1); break; }
        b = a; a = ref;
    }
}

// DFMA_OF: of every 4 warps, how many run the FP64 multiplication (0 = none, 4 = all)
template <int DFMA_OF>
__global__ void __launch_bounds__(128) k_mix(uint32_t* out, int iters) {
    const unsigned warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    if ((int)warp < DFMA_OF) {
        fq52::el a, b;
#pragma unroll
        for (int i = 0; i < 8; i++) { a.l[i] = (double)(1000 + i + threadIdx.x); b.l[i] = (double)(77 + 3 * i + blockIdx.x % 100); }
        for (int it = 0; it < iters; it++) { a = fq52::mul(a, b); b = fq52::mul(b, a); }
#pragma unroll
        for (int i = 0; i < 8; i++) acc ^= (uint32_t)(long long)(a.l[i] + b.l[i]);
    } else {
        fq_t a = fq_t::one(), b = fq_t::r2();
        a.l[0] ^= threadIdx.x; b.l[1] ^= blockIdx.x;
        for (int it = 0; it < iters; it++) { a = a * b; b = b * a; }
        const fq_t s = a + b;
#pragma unroll
        for (int i = 0; i < 12; i++) acc ^= s.l[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class K>
static float time_kernel(K kern, dim3 grid, dim3 block, uint32_t* out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) kern<<<grid, block>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        kern<<<grid, block>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

// correctness on the device: the FP64 product of two small values against the integer path
__global__ void k_check(int* bad) {
    fq52::el a, b;
    for (int i = 0; i < 8; i++) { a.l[i] = (double)(123456789 + 1000003ll * i * (threadIdx.x + 1)); b.l[i] = (double)(987654321 + 77ll * i + threadIdx.x); }
    a.l[7] = (double)(threadIdx.x & 0xffff); b.l[7] = 5.0;
    // a * b * R52^-1, then the same through 32-bit limbs: compare after converting a, b to 12 x u32
    auto to32 = [](const fq52::el& e) {
        fq_t r = fq_t::zero();
        // 8 x 52 bits -> 12 x 32 bits
        unsigned long long v[8];
        for (int i = 0; i < 8; i++) v[i] = (unsigned long long)e.l[i];
        for (int w = 0; w < 12; w++) {
            const int bit = 32 * w, li = bit / 52, off = bit % 52;
            unsigned long long x = v[li] >> off;
            if (off > 20 && li + 1 < 8) x |= v[li + 1] << (52 - off);
            r.l[w] = (uint32_t)x;
        }
        return r;
    };
    const fq52::el c52 = fq52::mul(a, b);
    const fq_t two32 = from_u64<FqParams>(1ull << 32);
    const fq_t ref = (to32(a) * to32(b)) * inverse(two32);
    if (!(to32(c52) == ref)) atomicAdd(bad, 1);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t* out;
    CK(cudaMalloc(&out, (size_t)sms * 32 * 128 * sizeof(uint32_t)));
    int* bad;
    CK(cudaMalloc(&bad, sizeof(int)));
    CK(cudaMemset(bad, 0, sizeof(int)));
    k_check<<<4, 128>>>(bad);
    int hbad = -1;
    CK(cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemset(bad, 0, sizeof(int)));
    k_check48<<<64, 128>>>(bad);
    int hbad48 = -1;
    CK(cudaMemcpy(&hbad48, bad, sizeof(int), cudaMemcpyDeviceToHost));
    const int it = 256;
    printf("{\"gpu\": \"%s\", \"device_check_mismatches\": %d, \"device_check48_mismatches\": %d", prop.name, hbad, hbad48);
    for (int bps : {2, 4, 6}) {
        dim3 grid(sms * bps), block(128);
        const double muls = (double)grid.x * 128 * it * 2;
        float ms;
        ms = time_kernel(k_mix<0>, grid, block, out, it);
        printf(", \"imad_G_per_s_%dwarps\": %.2f", bps * 4, muls / ms / 1e6);
        ms = time_kernel(k_mix<4>, grid, block, out, it);
        printf(", \"dfma_G_per_s_%dwarps\": %.2f", bps * 4, muls / ms / 1e6);
        ms = time_kernel(k_mix<2>, grid, block, out, it);
        printf(", \"hybrid_2of4_G_per_s_%dwarps\": %.2f", bps * 4, muls / ms / 1e6);
        ms = time_kernel(k_mix<1>, grid, block, out, it);
        printf(", \"hybrid_1of4_G_per_s_%dwarps\": %.2f", bps * 4, muls / ms / 1e6);
        ms = time_kernel(k_mix48<4>, grid, block, out, it);
        printf(", \"fp64only_G_per_s_%dwarps\": %.2f", bps * 4, muls / ms / 1e6);
        ms = time_kernel(k_mix48<2>, grid, block, out, it);
        printf(", \"hybrid48_2of4_G_per_s_%dwarps\": %.2f", bps * 4, muls / ms / 1e6);
        ms = time_kernel(k_mix48<1>, grid, block, out, it);
        printf(", \"hybrid48_1of4_G_per_s_%dwarps\": %.2f", bps * 4, muls / ms / 1e6);
        ms = time_kernel(k_mix48<3>, grid, block, out, it);
        printf(", \"hybrid48_3of4_G_per_s_%dwarps\": %.2f", bps * 4, muls / ms / 1e6);
    }
    CK(cudaGetLastError());
    printf("}\n");
    return 0;
}

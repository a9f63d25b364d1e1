// Integer-pipe microbenchmark for B200: measures the peak the MSM / NTT kernels are
// reported against (MEASURED_PEAKS.json has no integer figure).
//   * mad.lo.u32 / mad.hi.u32 / mad.wide.u32 throughput, 8 independent chains per thread
//   * Fr (8-limb) and Fq (12-limb) Montgomery multiplications per second as implemented
//     in csrc/arith.cuh (dependent chain per thread, all SMs, several occupancies)
// Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../dusk-plonk_b200/csrc/arith.cuh"
using namespace zkp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void k_imad(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[8];
    uint64_t w[8];
    uint32_t b = seed | 1, c = threadIdx.x * 2654435761u + 12345u;
    uint64_t c64 = ((uint64_t)c << 32) | seed;
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = c + i * 977u; w[i] = ((uint64_t)a[i] << 32) | i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 2) asm volatile("{ .reg .u32 t, u; mov.b64 {t, u}, %0; xor.b32 t, t, u; mad.wide.u32 %0, t, %1, %2; }" : "+l"(w[i]) : "r"(b), "l"(c64));
            if (MODE == 3) {  // lo+hi pair with carry, as the Montgomery rows issue them
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
                             : "+r"(a[i]), "+r"(a[(i + 1) & 7]) : "r"(c), "r"(b));
            }
            if (MODE == 4) asm volatile("mad.lo.u32 %0, %0, 0x53bda402, %1;" : "+r"(a[i]) : "r"(c));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class F>
__global__ void k_fmul(F* out, int iters) {
    F a = F::one(), b = F::r2();
    a.l[0] ^= threadIdx.x; b.l[1] ^= blockIdx.x;
    for (int it = 0; it < iters; it++) { a = a * b; b = b * a; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
}

template <class F>
__global__ void k_fmul_ilp2(F* out, int iters) {
    F a = F::one(), b = F::r2(), c = F::r2(), d = F::one();
    a.l[0] ^= threadIdx.x; b.l[1] ^= blockIdx.x; c.l[2] ^= threadIdx.x; d.l[3] ^= blockIdx.x;
    for (int it = 0; it < iters; it++) { a = a * b; c = c * d; b = b * a; d = d * c; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}

template <class K, class... A>
static float time_kernel(K kern, dim3 grid, dim3 block, A... args) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) kern<<<grid, block>>>(args...);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        kern<<<grid, block>>>(args...);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t* out;
    CK(cudaMalloc(&out, (size_t)sms * 16 * 1024 * 64));
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    const int iters = 4096;
    const char* names[5] = {"mad_lo", "mad_hi", "mad_wide", "mad_lo_hi_pair", "mad_lo_imm"};
    for (int mode = 0; mode < 5; mode++) {
        dim3 grid(sms * 8), block(256);
        float ms = 0;
        if (mode == 0) ms = time_kernel(k_imad<0>, grid, block, out, 3u, iters);
        if (mode == 1) ms = time_kernel(k_imad<1>, grid, block, out, 3u, iters);
        if (mode == 2) ms = time_kernel(k_imad<2>, grid, block, out, 3u, iters);
        if (mode == 3) ms = time_kernel(k_imad<3>, grid, block, out, 3u, iters);
        if (mode == 4) ms = time_kernel(k_imad<4>, grid, block, out, 3u, iters);
        double ops = (double)sms * 8 * 256 * iters * 8 * (mode == 3 ? 2 : 1);
        printf(", \"%s_Tops\": %.3f", names[mode], ops / ms / 1e9);
    }
    for (int tpb_i = 0; tpb_i < 3; tpb_i++) {
        const int blocks_per_sm[3] = {2, 4, 8};
        dim3 grid(sms * blocks_per_sm[tpb_i]), block(128);
        const int it = 512;
        float ms = time_kernel(k_fmul<fr_t>, grid, block, (fr_t*)out, it);
        printf(", \"fr_mul_G_per_s_%dwarps\": %.2f", blocks_per_sm[tpb_i] * 4, (double)grid.x * 128 * it * 2 / ms / 1e6);
        ms = time_kernel(k_fmul<fq_t>, grid, block, (fq_t*)out, it);
        printf(", \"fq_mul_G_per_s_%dwarps\": %.2f", blocks_per_sm[tpb_i] * 4, (double)grid.x * 128 * it * 2 / ms / 1e6);
        ms = time_kernel(k_fmul_ilp2<fq_t>, grid, block, (fq_t*)out, it);
        printf(", \"fq_mul_ilp2_G_per_s_%dwarps\": %.2f", blocks_per_sm[tpb_i] * 4, (double)grid.x * 128 * it * 4 / ms / 1e6);
    }
    printf("}\n");
    return 0;
}

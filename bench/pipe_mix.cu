// Pipe-mix microbenchmark for sm_100a: how much multiply throughput sits beside the IMAD pipe.
// Input to the round-2 design question of DESIGN.md section 4 (moving part of the Fq multiplication to
// the FP64 pipe: a 52-bit-limb product = 2 DFMA + 1 DADD, Emmart's floating-point big-integer scheme).
// Measures, with 8 independent dependency chains per thread and 8 x 256 threads per SM:
//   dfma        fma.rz.f64 alone
//   imad_wide   mad.wide.u32 alone (the instruction the field arithmetic is made of)
//   dfma+imad   both in the same loop (one of each per chain step): do the pipes add up?
//   imad+iadd3  mad.wide.u32 with a 3-input add per step (the ALU pipe, for Karatsuba-style trades)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void k_mix(double* out, double seed, int iters) {
    double d[8];
    uint64_t w[8];
    uint32_t s[8];
    const double b = seed * 1.0000001, c = seed + threadIdx.x;
    const uint32_t bi = (uint32_t)threadIdx.x * 2654435761u | 1u;
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = c + i; w[i] = ((uint64_t)(bi + i) << 32) | i; s[i] = bi + 7 * i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(b), "d"(c));
            if (MODE == 1 || MODE == 2 || MODE == 3)
                asm volatile("{ .reg .u32 t, u; mov.b64 {t, u}, %0; mad.wide.u32 %0, t, %1, %0; }" : "+l"(w[i]) : "r"(bi));
            if (MODE == 3) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(s[i]) : "r"(bi), "r"(s[(i + 1) & 7]));
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += d[i] + (double)w[i] + (double)s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K>
static float time_kernel(K kern, dim3 grid, dim3 block, double* out, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<grid, block>>>(out, 1.5, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    kern<<<grid, block>>>(out, 1.5, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, iters = 4096;
    double* out;
    CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(double)));
    dim3 grid(sms * 8), block(256);
    const double steps = (double)sms * 8 * 256 * iters * 8;   // chain steps executed per launch
    const float t0 = time_kernel(k_mix<0>, grid, block, out, iters);
    const float t1 = time_kernel(k_mix<1>, grid, block, out, iters);
    const float t2 = time_kernel(k_mix<2>, grid, block, out, iters);
    const float t3 = time_kernel(k_mix<3>, grid, block, out, iters);
    CK(cudaGetLastError());
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_Tops\": %.3f, \"imad_wide_Tops\": %.3f, "
           "\"dfma_plus_imad_wide\": {\"steps_Tops\": %.3f, \"note\": \"one DFMA and one IMAD.WIDE per step\"}, "
           "\"imad_wide_plus_2add\": {\"steps_Tops\": %.3f, \"note\": \"one IMAD.WIDE and two 32-bit adds per step\"}}\n",
           prop.name, sms, steps / t0 / 1e9, steps / t1 / 1e9, steps / t2 / 1e9, steps / t3 / 1e9);
    return 0;
}

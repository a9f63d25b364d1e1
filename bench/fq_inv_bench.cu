// Throughput of the divergence-free binary-GCD Fq inversion (bench/fq_inv32.cuh) against the Fermat chain of
// arith.cuh on sm_100a, in units of Fq multiplications.  EXPERIMENT / groundwork for round 2.
#include <cuda_runtime.h>
#include <stdio.h>

#include "../dusk-plonk_b200/csrc/arith.cuh"
#include "../dusk-plonk_b200/csrc/fq_inv32.cuh"

using namespace zkp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ fq_t modulus() { return fq_t::modulus(); }

// Montgomery-domain inverse through the binary GCD: (a R)^-1 by inv_mod, times R^3 R^-1
__device__ __forceinline__ fq_t inverse_bingcd(const fq_t& a, const fq_t& r3) {
    const fq_t m = modulus();
    fq_t x;
    fqinv::inv_mod<12>(a.l, m.l, 0x7ffdu /* -p^-1 mod 2^15 */, 51, x.l);
    return x * r3;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_inv(fq_t* out, int iters) {
    fq_t a = fq_t::r2();
    a.l[0] ^= threadIdx.x * 2654435761u; a.l[2] ^= blockIdx.x * 40503u;
    a.l[11] &= 0x0fffffffu;
    const fq_t r2 = fq_t::r2();
    const fq_t r3 = r2 * r2;
    const fq_t one = fq_t::one();
    for (int it = 0; it < iters; it++) {
        a = (MODE == 0 ? inverse(a) : inverse_bingcd(a, r3)) + one;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

__global__ void k_mul(fq_t* out, int iters) {
    fq_t a = fq_t::one(), b = fq_t::r2();
    a.l[0] ^= threadIdx.x; b.l[1] ^= blockIdx.x;
    for (int it = 0; it < iters; it++) { a = a * b; b = b * a; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
}

__global__ void k_check(int* bad) {
    fq_t a = fq_t::r2();
    a.l[0] ^= threadIdx.x * 2654435761u; a.l[5] ^= blockIdx.x * 977u;
    a.l[11] &= 0x0fffffffu;
    const fq_t r2 = fq_t::r2();
    const fq_t r3 = r2 * r2;
    for (int it = 0; it < 4; it++) {
        const fq_t f = inverse(a), g = inverse_bingcd(a, r3);
        if (!(f == g)) atomicAdd(bad, 1);
        a = f + fq_t::one();
    }
}

template <class K, class... A>
static float time_kernel(K kern, dim3 grid, dim3 block, A... args) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, block>>>(args...);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
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
    fq_t* out;
    CK(cudaMalloc(&out, (size_t)sms * 8 * 128 * sizeof(fq_t)));
    int* bad;
    CK(cudaMalloc(&bad, sizeof(int)));
    CK(cudaMemset(bad, 0, sizeof(int)));
    k_check<<<32, 128>>>(bad);
    int hbad = -1;
    CK(cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost));
    dim3 grid(sms * 4), block(128);
    const double threads = (double)grid.x * 128;
    const float t_mul = time_kernel(k_mul, grid, block, out, 512);
    const float t_fermat = time_kernel(k_inv<0>, grid, block, out, 8);
    const float t_gcd = time_kernel(k_inv<1>, grid, block, out, 8);
    CK(cudaGetLastError());
    const double mul_per_s = threads * 512 * 2 / (t_mul * 1e-3);
    const double fermat_per_s = threads * 8 / (t_fermat * 1e-3), gcd_per_s = threads * 8 / (t_gcd * 1e-3);
    printf("{\"gpu\": \"%s\", \"device_check_mismatches\": %d, \"fq_mul_G_per_s\": %.2f, \"fermat_inv_M_per_s\": %.1f, "
           "\"bingcd_inv_M_per_s\": %.1f, \"fermat_inv_in_muls\": %.1f, \"bingcd_inv_in_muls\": %.1f}\n",
           prop.name, hbad, mul_per_s / 1e9, fermat_per_s / 1e6, gcd_per_s / 1e6, mul_per_s / fermat_per_s,
           mul_per_s / gcd_per_s);
    return 0;
}

// exact_chain_probe.cu — are the shared-reciprocal / unchecked sequences of tr_device_math.cuh (xunit3_mid, xnormalize3_mid)
// bit-identical to the library's IEEE operations (__fdiv_rn, __fsqrt_rn) they replace?  Counts mismatching results over
// random inputs.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../transmission_renderer_b200/csrc exact_chain_probe.cu
#include <cstdio>
#include <cstdint>
#include "tr_device_math.cuh"
using namespace trd;

__device__ uint32_t rng(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
__device__ float uni(uint64_t& s, float lo, float hi) { return lo + (hi - lo) * (rng(s) * (1.0f / 2147483648.0f)); }

__global__ void probe(unsigned long long* out, int iters) {
    uint64_t s = 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long bad_unit = 0, bad_norm = 0, bad_sqrt = 0;
    for (int i = 0; i < iters; i++) {
        const float scale = exp2f(uni(s, -8.0f, 8.0f));
        const f3 a = mk3(uni(s, -1, 1) * scale, uni(s, -1, 1) * scale, uni(s, -1, 1) * scale);
        const f3 u0 = xdivs3(a, xsqrt(xdot3(a, a))), u1 = xunit3_mid(a);
        bad_unit += (__float_as_uint(u0.x) != __float_as_uint(u1.x)) + (__float_as_uint(u0.y) != __float_as_uint(u1.y)) +
                    (__float_as_uint(u0.z) != __float_as_uint(u1.z));
        const f3 n0 = xnormalize3(a), n1 = xnormalize3_mid(a);
        bad_norm += (__float_as_uint(n0.x) != __float_as_uint(n1.x)) + (__float_as_uint(n0.y) != __float_as_uint(n1.y)) +
                    (__float_as_uint(n0.z) != __float_as_uint(n1.z));
        const float x = xdot3(a, a);
        bad_sqrt += __float_as_uint(xsqrt(x)) != __float_as_uint(xsqrt_mid(x));
    }
    atomicAdd(out, bad_unit); atomicAdd(out + 1, bad_norm); atomicAdd(out + 2, bad_sqrt);
}

int main() {
    unsigned long long* d; unsigned long long h[3] = {0, 0, 0};
    cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
    const int iters = 2000, blocks = 1024, threads = 256;
    probe<<<blocks, threads>>>(d, iters);
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    const double n = (double)iters * blocks * threads;
    printf("vectors tested %.0f: unit-vector components differing %llu, normalize components differing %llu, sqrt differing %llu\n", n, h[0], h[1], h[2]);
    return cudaGetLastError() != cudaSuccess;
}

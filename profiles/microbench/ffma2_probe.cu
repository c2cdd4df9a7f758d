// ffma2_probe.cu — does the packed FP32 instruction of sm_100 (FFMA2, PTX fma.rn.f32x2) free issue slots on B200?
// Four kernels with the same number of FP32 FMAs per thread:
//   scalar      : 8 independent FFMA chains
//   packed      : 4 independent FFMA2 chains (same FMAs, half the instructions)
//   scalar+alu  : the scalar kernel with one integer LOP3/IADD per FFMA interleaved
//   packed+alu  : the packed kernel with the same integer work
// If FFMA2 runs at half the instruction rate of FFMA (same lanes), `packed` matches `scalar`, and `packed+alu`
// stays near it while `scalar+alu` halves: the integer work rides in the freed issue slots.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ITERS = 4096;

template <bool PACKED, bool ALU>
__global__ void __launch_bounds__(256) probe(float* out, unsigned* iout, float a, float b) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned x0 = tid, x1 = tid * 3u + 1u, x2 = tid ^ 0x5555u, x3 = tid + 77u;
    if (PACKED) {
        float2 c0 = make_float2(tid * 1e-9f, 1.0f), c1 = make_float2(2.0f, 3.0f), c2 = make_float2(4.0f, 5.0f), c3 = make_float2(6.0f, 7.0f);
        const float2 aa = make_float2(a, a), bb = make_float2(b, b);
#pragma unroll 8
        for (int i = 0; i < ITERS; i++) {
            c0 = __ffma2_rn(c0, aa, bb);
            if (ALU) { x0 = (x0 ^ x1) + 0x9e37u; x1 = (x1 & x2) + x0; }
            c1 = __ffma2_rn(c1, aa, bb);
            if (ALU) { x2 = (x2 | x3) + 0x7f4au; x3 = (x3 ^ x0) + x2; }
            c2 = __ffma2_rn(c2, aa, bb);
            if (ALU) { x0 = (x0 ^ x2) + 0x1234u; x1 = (x1 & x3) + x0; }
            c3 = __ffma2_rn(c3, aa, bb);
            if (ALU) { x2 = (x2 | x1) + 0x4321u; x3 = (x3 ^ x1) + x2; }
        }
        out[tid] = c0.x + c0.y + c1.x + c1.y + c2.x + c2.y + c3.x + c3.y;
    } else {
        float c0 = tid * 1e-9f, c1 = 1.0f, c2 = 2.0f, c3 = 3.0f, c4 = 4.0f, c5 = 5.0f, c6 = 6.0f, c7 = 7.0f;
#pragma unroll 8
        for (int i = 0; i < ITERS; i++) {
            c0 = fmaf(c0, a, b); c1 = fmaf(c1, a, b);
            if (ALU) { x0 = (x0 ^ x1) + 0x9e37u; x1 = (x1 & x2) + x0; }
            c2 = fmaf(c2, a, b); c3 = fmaf(c3, a, b);
            if (ALU) { x2 = (x2 | x3) + 0x7f4au; x3 = (x3 ^ x0) + x2; }
            c4 = fmaf(c4, a, b); c5 = fmaf(c5, a, b);
            if (ALU) { x0 = (x0 ^ x2) + 0x1234u; x1 = (x1 & x3) + x0; }
            c6 = fmaf(c6, a, b); c7 = fmaf(c7, a, b);
            if (ALU) { x2 = (x2 | x1) + 0x4321u; x3 = (x3 ^ x1) + x2; }
        }
        out[tid] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
    }
    if (ALU) iout[tid] = x0 + x1 + x2 + x3;
}

template <bool PACKED, bool ALU>
float run(const char* name, float* out, unsigned* iout, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) probe<PACKED, ALU><<<blocks, 256>>>(out, iout, 0.999f, 0.001f);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; r++) probe<PACKED, ALU><<<blocks, 256>>>(out, iout, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double flops = 2.0 * 8.0 * ITERS * 256.0 * blocks;
    printf("%-12s %8.3f ms  %7.2f TFLOP/s (fp32 FMA)\n", name, ms, flops / (ms * 1e-3) / 1e12);
    return ms;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8 * 4;
    float* out;
    unsigned* iout;
    cudaMalloc(&out, sizeof(float) * 256 * blocks);
    cudaMalloc(&iout, sizeof(unsigned) * 256 * blocks);
    printf("SMs %d\n", sms);
    run<false, false>("scalar", out, iout, blocks);
    run<true, false>("packed", out, iout, blocks);
    run<false, true>("scalar+alu", out, iout, blocks);
    run<true, true>("packed+alu", out, iout, blocks);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

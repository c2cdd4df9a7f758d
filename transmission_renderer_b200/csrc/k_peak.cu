// k_peak.cu — diagnostics: measured roofline denominators for this GPU, taken in the same process
// as the benchmark (MEASURED_PEAKS.json has no FP32 figure).
//   tr_measure_fp32_peak  dependent-FFMA chains, 8 independent accumulators per thread, all SMs
//   tr_measure_hbm_copy   STREAM-style copy (read + write bytes), 128-bit accesses
#include "tr_internal.h"

namespace {

__global__ void __launch_bounds__(256) ffma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace

using namespace tr;

extern "C" {

int32_t tr_measure_fp32_peak(tr_ctx* c, float* tflops) {
    if (!c || !tflops) return fail(TR_ERR_INVALID_ARG, "tr_measure_fp32_peak: null");
    TR_CUDA(cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
    DevBuf out;
    TR_TRY(out.ensure((size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    TR_CUDA(cudaEventCreate(&e0));
    TR_CUDA(cudaEventCreate(&e1));
    float best = 0.0f;
    for (int rep = 0; rep < 5; rep++) {
        TR_CUDA(cudaEventRecord(e0, c->stream));
        ffma_kernel<<<blocks, threads, 0, c->stream>>>(out.as<float>(), iters, 0.999f, 0.001f);
        TR_CUDA(cudaEventRecord(e1, c->stream));
        TR_CUDA(cudaEventSynchronize(e1));
        float ms = 0.0f;
        TR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    out.release();
    *tflops = best;
    return TR_OK;
}

int32_t tr_measure_hbm_copy(tr_ctx* c, float* gbs) {
    if (!c || !gbs) return fail(TR_ERR_INVALID_ARG, "tr_measure_hbm_copy: null");
    TR_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)1 << 30;
    DevBuf a, b;
    TR_TRY(a.ensure(bytes));
    TR_TRY(b.ensure(bytes));
    TR_CUDA(cudaMemsetAsync(a.p, 1, bytes, c->stream));
    cudaEvent_t e0, e1;
    TR_CUDA(cudaEventCreate(&e0));
    TR_CUDA(cudaEventCreate(&e1));
    float best = 0.0f;
    for (int rep = 0; rep < 6; rep++) {
        TR_CUDA(cudaEventRecord(e0, c->stream));
        copy_kernel<<<c->sm_count * 16, 256, 0, c->stream>>>(a.as<uint4>(), b.as<uint4>(), bytes / 16);
        TR_CUDA(cudaEventRecord(e1, c->stream));
        TR_CUDA(cudaEventSynchronize(e1));
        float ms = 0.0f;
        TR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const float g = (float)(2.0 * bytes / (ms * 1e-3) / 1e9);
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    a.release();
    b.release();
    *gbs = best;
    return TR_OK;
}

}  // extern "C"

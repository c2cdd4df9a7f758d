// k_tonemap.cu — K7: Lottes tonemap + sRGB8 encode (sm_100a).
// Reference: shader/src/tonemapping.rs:9-26, fragment_tonemap shader/src/lib.rs:683-697,
// swapchain format B8G8R8A8_SRGB src/main.rs:175 (the sRGB OETF + UNORM8 store was fixed function).
// Black pixels: divide by max(max_element, FLT_MIN); inputs are clamped to [0, 65504] and NaN -> 0
// (fp16 overflow / NaN have no defined result in the reference; see oracle/tonemap.c).
// 8 B read + 4 B written per pixel; 4 pixels per thread, 128-bit loads and stores.
#include <float.h>

#include "tr_internal.h"

using namespace trd;

namespace {

// The pass is bound by the MUFU pipe, not by HBM (ncu: XU 83 %): every pow is an lg2 and an ex2, and the straightforward
// transcription of tonemapping.rs:9-26 + the sRGB OETF needs 26 of them per pixel.  Everything between the first logarithm and
// the last exponential is a product of powers, so it is carried in log2 space: out_k = (c_k / max)^e0 -> lerp -> ^cross_saturation
// * tonemapped_max -> OETF needs lg2(c_k), one ex2 and one lg2 around the lerp (a sum, which has no log form) and one ex2 for the
// OETF — 4 per channel instead of 6, 18 per pixel.  Same mathematical function; against the CPU oracle's powf chain the sRGB8
// bytes differ by at most one code at a rounding boundary (tolerance 2, tests/test_gpu_parity.py).
__device__ __forceinline__ float lg2(float x) { return __log2f(x); }     // lg2(0) = -inf, which every use below wants
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }   // ex2(-inf) = 0

__device__ __forceinline__ uint32_t srgb8_log(float l) {   // l = lg2 of the linear value, already <= 0 (value <= 1)
    const float s = l <= -8.3192835f /* lg2(0.0031308) */ ? ex2(l) * 12.92f : fmaf(1.055f, ex2(l * (1.0f / 2.4f)), -0.055f);
    return (uint32_t)floorf(fmaf(s, 255.0f, 0.5f));
}

__device__ __forceinline__ uint32_t tonemap_px(uint2 v, const tr_baked_lottes_tonemapper_params& p) {
    f4 c = unpack_rgba16f(v);
    c.x = c.x > 0.0f ? fminf(c.x, 65504.0f) : 0.0f;
    c.y = c.y > 0.0f ? fminf(c.y, 65504.0f) : 0.0f;
    c.z = c.z > 0.0f ? fminf(c.z, 65504.0f) : 0.0f;
    const float mx = fmaxf(fmaxf(c.x, fmaxf(c.y, c.z)), FLT_MIN);
    const float lmx = lg2(mx);
    const float lz = p.a * lmx;                                  // tonemap_inner, tonemapping.rs:9-12: z = max^a
    const float tm = ex2(lz) / fmaf(ex2(lz * p.d), p.b, p.c);    // z / (z^d b + c)
    const float ltm = lg2(tm);
    const float e0 = p.saturation / p.cross_saturation;
    const float cross = ex2(ltm * p.crosstalk);
    const float in[3] = {c.x, c.y, c.z};
    uint32_t out = 0xff000000u;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float r = ex2((lg2(in[k]) - lmx) * e0);                  // (c_k / max)^e0; a zero channel gives ex2(-inf) = 0
        r = fmaf(1.0f - r, cross, r);                            // ratio.lerp(ONE, tm^crosstalk)
        const float l = fminf(fmaf(lg2(r), p.cross_saturation, ltm), 0.0f);   // lg2(min(r^cross_saturation tm, 1))
        out |= srgb8_log(l == l ? l : -INFINITY) << (8 * k);     // -inf + inf (r = inf cannot happen; tm = 0 with r = 0 can): black
    }
    return out;
}

__global__ void __launch_bounds__(256) tonemap_kernel(const uint2* __restrict__ hdr, uint32_t* __restrict__ out,
                                                      uint32_t px_begin, uint32_t px_end,
                                                      tr_baked_lottes_tonemapper_params p) {
    const uint32_t n = px_end - px_begin;
    const uint32_t stride = gridDim.x * blockDim.x * 4;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        const uint32_t g = px_begin + i;
        if ((g & 3u) == 0u && i + 4 <= n) {
            const uint4 a = __ldcs(reinterpret_cast<const uint4*>(hdr + g));
            const uint4 b = __ldcs(reinterpret_cast<const uint4*>(hdr + g + 2));
            uint4 o;
            o.x = tonemap_px(make_uint2(a.x, a.y), p);
            o.y = tonemap_px(make_uint2(a.z, a.w), p);
            o.z = tonemap_px(make_uint2(b.x, b.y), p);
            o.w = tonemap_px(make_uint2(b.z, b.w), p);
            *reinterpret_cast<uint4*>(out + g) = o;
        } else {
            for (uint32_t k = 0; k < 4 && i + k < n; k++) out[g + k] = tonemap_px(hdr[g + k], p);
        }
    }
}

}  // namespace

namespace tr {
int32_t launch_tonemap(const uint2* hdr, uchar4* out, uint32_t px_begin, uint32_t px_end,
                       const tr_baked_lottes_tonemapper_params& params, int sm_count, cudaStream_t s) {
    if (px_end <= px_begin) return TR_OK;
    const uint32_t n = px_end - px_begin;
    uint32_t blocks = (n / 4 + 255) / 256;
    const uint32_t cap = (uint32_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    tonemap_kernel<<<blocks, 256, 0, s>>>(hdr, reinterpret_cast<uint32_t*>(out), px_begin, px_end, params);
    count_launches(1);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}
}  // namespace tr

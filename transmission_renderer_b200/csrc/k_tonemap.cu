// k_tonemap.cu — K7: Lottes tonemap + sRGB8 encode (sm_100a).
// Reference: shader/src/tonemapping.rs:9-26, fragment_tonemap shader/src/lib.rs:683-697,
// swapchain format B8G8R8A8_SRGB src/main.rs:175 (the sRGB OETF + UNORM8 store was fixed function).
// Black pixels: divide by max(max_element, FLT_MIN); inputs are clamped to [0, 65504] and NaN -> 0
// (fp16 overflow / NaN have no defined result in the reference; see oracle/tonemap.c).
// HBM-bound: 8 B read + 4 B written per pixel; 4 pixels per thread, 128-bit loads and stores.
#include <float.h>

#include "tr_internal.h"

using namespace trd;

namespace {

__device__ __forceinline__ float fpow(float x, float y) { return exp2f(y * __log2f(x)); }

__device__ __forceinline__ uint32_t srgb8(float c) {
    c = fminf(fmaxf(c, 0.0f), 1.0f);
    float s = c <= 0.0031308f ? c * 12.92f : fmaf(1.055f, fpow(c, 1.0f / 2.4f), -0.055f);
    return (uint32_t)floorf(fmaf(s, 255.0f, 0.5f));
}

__device__ __forceinline__ uint32_t tonemap_px(uint2 v, const tr_baked_lottes_tonemapper_params& p) {
    f4 c = unpack_rgba16f(v);
    c.x = c.x > 0.0f ? fminf(c.x, 65504.0f) : 0.0f;
    c.y = c.y > 0.0f ? fminf(c.y, 65504.0f) : 0.0f;
    c.z = c.z > 0.0f ? fminf(c.z, 65504.0f) : 0.0f;
    float mx = fmaxf(fmaxf(c.x, fmaxf(c.y, c.z)), FLT_MIN);
    float inv = 1.0f / mx;
    float z = fpow(mx, p.a);                                  // tonemap_inner, tonemapping.rs:9-12
    float tm = z / fmaf(fpow(z, p.d), p.b, p.c);
    float e0 = p.saturation / p.cross_saturation;
    float cross = fpow(tm, p.crosstalk);
    float out[3] = {c.x * inv, c.y * inv, c.z * inv};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float r = fpow(out[k], e0);
        r = fmaf(1.0f - r, cross, r);                          // ratio.lerp(ONE, tm^crosstalk)
        r = fpow(r, p.cross_saturation);
        out[k] = fminf(fmaxf(r * tm, 0.0f), 1.0f);
    }
    return srgb8(out[0]) | (srgb8(out[1]) << 8) | (srgb8(out[2]) << 16) | 0xff000000u;
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

/* oracle/mips.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The opaque-frame mip chain.  Reference call site: src/main.rs:2054-2063
 * (`generate_mips`), level count src/main.rs:2590-2592.  The body lives in
 * ash-opinionated-abstractions @ 8591c309 (Cargo.lock:70-72, NOT in
 * /root/reference): a chain of vkCmdBlitImage(LINEAR) level i -> i+1 on an
 * R16G16B16A16_SFLOAT image.  Restated here from the Vulkan blit rules
 * (SURVEY.md Appendix E):
 *   size(k+1) = max(1, floor(size(k)/2))
 *   dst texel (x,y) samples level k bilinearly at
 *       ((x+0.5) * w_k/w_{k+1}, (y+0.5) * h_k/h_{k+1})   (texel space,
 *       centres at +0.5, clamp-to-edge), lerp x then y in fp32,
 *   one round-to-nearest-even to fp16 per level.
 * For even source sizes the fractions are exactly 0.5 (2x2 box).
 * PARITY: NOT PINNABLE at this boundary (vkCmdBlitImage: hardware behaviour in the reference; the filter is ours, oracle.h).
 */
#include "oracle.h"

/* src/main.rs:2590-2592: (min(w,h) as f32).log2() as u32 + 1 */
uint32_t orc_mip_levels_for_size(uint32_t w, uint32_t h) {
    uint32_t m = w < h ? w : h;
    float l = log2f((float)m);
    return (uint32_t)l + 1u;
}

void orc_mip_size(uint32_t w, uint32_t h, uint32_t level, uint32_t* lw, uint32_t* lh) {
    for (uint32_t i = 0; i < level; i++) {
        w = w / 2 > 1 ? w / 2 : 1;
        h = h / 2 > 1 ? h / 2 : 1;
    }
    *lw = w;
    *lh = h;
}

static void axis_setup(uint32_t d, uint32_t src_size, uint32_t dst_size, uint32_t* i0, uint32_t* i1, float* frac) {
    float scale = (float)src_size / (float)dst_size;
    float p = ((float)d + 0.5f) * scale - 0.5f;
    float fl = floorf(p);
    int64_t i = (int64_t)fl;
    int64_t hi = (int64_t)src_size - 1;
    int64_t a = i < 0 ? 0 : (i > hi ? hi : i);
    int64_t b = i + 1 < 0 ? 0 : (i + 1 > hi ? hi : i + 1);
    *i0 = (uint32_t)a;
    *i1 = (uint32_t)b;
    *frac = p - fl;
}

void orc_downsample_level(const uint16_t* src, uint32_t sw, uint32_t sh, uint16_t* dst, uint32_t dw, uint32_t dh) {
#pragma omp parallel for schedule(static)
    for (int64_t yy = 0; yy < (int64_t)dh; yy++) {
        uint32_t y = (uint32_t)yy;
        uint32_t y0, y1;
        float fy;
        axis_setup(y, sh, dh, &y0, &y1, &fy);
        for (uint32_t x = 0; x < dw; x++) {
            uint32_t x0, x1;
            float fx;
            axis_setup(x, sw, dw, &x0, &x1, &fx);
            for (int c = 0; c < 4; c++) {
                float t00 = f16_bits_to_f32(src[((size_t)y0 * sw + x0) * 4 + c]);
                float t10 = f16_bits_to_f32(src[((size_t)y0 * sw + x1) * 4 + c]);
                float t01 = f16_bits_to_f32(src[((size_t)y1 * sw + x0) * 4 + c]);
                float t11 = f16_bits_to_f32(src[((size_t)y1 * sw + x1) * 4 + c]);
                float top = t00 + (t10 - t00) * fx;
                float bot = t01 + (t11 - t01) * fx;
                dst[((size_t)y * dw + x) * 4 + c] = f32_to_f16_bits(top + (bot - top) * fy);
            }
        }
    }
}

/* array wrappers so the tests can pin the fp16 conversions against IEEE (numpy) */
void orc_f32_to_f16(const float* in, uint16_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = f32_to_f16_bits(in[i]);
}
void orc_f16_to_f32(const uint16_t* in, float* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = f16_bits_to_f32(in[i]);
}

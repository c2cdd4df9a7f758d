/* oracle/tonemap.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of shader/src/tonemapping.rs:9-26 (LottesTonemapper) and
 * shader/src/lib.rs:683-697 (fragment_tonemap), followed by the sRGB8 store
 * of the swapchain format B8G8R8A8_SRGB (src/main.rs:175) which the Vulkan
 * implementation performed in the reference.
 * The baked parameters come from colstodian @ f2fb0f55 (not in
 * /root/reference); they are caller input here.
 *
 * Defined behaviour for black pixels: the reference divides by
 * max_element(color) (tonemapping.rs:16-17), 0/0 for a black pixel; we divide
 * by max(max_element, FLT_MIN) so black stays black.
 * Defined behaviour for non-finite / negative input: an RGBA16F store overflows to +inf above
 * 65504 and the reference's formula then evaluates inf/inf; we clamp each input component to
 * [0, 65504] and map NaN to 0 before tonemapping.
 * PARITY: pinned to the reference's compiled fragment_tonemap.spv + this sRGB encode (sRGB8 equal on 9 000 pixels, black
 * excluded: 0/0 in the reference); the baked constants are caller input (oracle.h, tests/test_reference_spirv.py).
 */
#include "oracle.h"

#include <float.h>

/* tonemapping.rs:9-12 */
static float tonemap_inner(float x, const tr_baked_lottes_tonemapper_params* p) {
    float z = powf(x, p->a);
    return z / (powf(z, p->d) * p->b + p->c);
}

static float sanitize(float x) {
    if (!(x > 0.0f)) return 0.0f; /* negative, -0, NaN */
    return x > 65504.0f ? 65504.0f : x;
}

/* tonemapping.rs:14-25 */
v3 orc_lottes_tonemap(v3 color, const tr_baked_lottes_tonemapper_params* p) {
    color = v3_new(sanitize(color.x), sanitize(color.y), sanitize(color.z));
    float max = f_max(v3_max_element(color), FLT_MIN);
    v3 ratio = v3_divs(color, max);
    float tonemapped_max = tonemap_inner(max, p);

    float e0 = p->saturation / p->cross_saturation;
    ratio = v3_new(powf(ratio.x, e0), powf(ratio.y, e0), powf(ratio.z, e0));
    ratio = v3_lerp(ratio, v3_splat(1.0f), powf(tonemapped_max, p->crosstalk));
    ratio = v3_new(powf(ratio.x, p->cross_saturation), powf(ratio.y, p->cross_saturation),
                   powf(ratio.z, p->cross_saturation));

    return v3_max(v3_min(v3_scale(ratio, tonemapped_max), v3_splat(1.0f)), v3_splat(0.0f));
}

/* sRGB OETF + UNORM8 store */
uint8_t orc_srgb8_encode(float c) {
    if (!(c > 0.0f)) c = 0.0f;
    if (c > 1.0f) c = 1.0f;
    float s = c <= 0.0031308f ? c * 12.92f : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
    return (uint8_t)floorf(s * 255.0f + 0.5f);
}

void orc_tonemap_frame(const uint16_t* hdr_f16, uint32_t w, uint32_t h, uint32_t y0, uint32_t y1,
                       const tr_baked_lottes_tonemapper_params* p, uint8_t* rgba8) {
    (void)h;
#pragma omp parallel for schedule(static)
    for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
        for (uint32_t x = 0; x < w; x++) {
            size_t i = (size_t)yy * w + x;
            v3 c = v3_new(f16_bits_to_f32(hdr_f16[i * 4]), f16_bits_to_f32(hdr_f16[i * 4 + 1]),
                          f16_bits_to_f32(hdr_f16[i * 4 + 2]));
            v3 t = orc_lottes_tonemap(c, p);
            rgba8[i * 4] = orc_srgb8_encode(t.x);
            rgba8[i * 4 + 1] = orc_srgb8_encode(t.y);
            rgba8[i * 4 + 2] = orc_srgb8_encode(t.z);
            rgba8[i * 4 + 3] = 255;
        }
    }
}

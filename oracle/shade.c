/* oracle/shade.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of the fragment stage:
 *   shader/src/lib.rs:37-162   fragment_transmission
 *   shader/src/lib.rs:164-249  fragment
 *   shader/src/lighting.rs:13-95, 127-220, 222-241, 261-313
 *   shared-structs/src/lib.rs:43-67, 90-138
 * plus the sampling rules the Vulkan implementation supplied in the reference
 *  and the G-buffer decode that stands in for the
 * rasteriser's varying interpolation, including the texture-mapped branches
 * (`textures.* != -1`: lighting.rs:222-313, lib.rs:66-77,120-124,190-194) with a
 * software restatement of the repeat sampler and implicit level of detail.
 * PARITY: pinned to the reference's compiled fragment.spv, fragment_transmission.spv, vertex_instanced_with_scale.spv and
 * depth_pre_pass_alpha_clip.spv: bit-equal fp32 pixels on whole frames (with a normal map <= 8e-6: the shipped module is
 * stale there).  The sampling rules themselves (what the Vulkan implementation supplied) are OURS to define
 * and are shared with the harness that runs the modules (oracle.h, tests/test_reference_spirv.py).
 */
#include "oracle.h"

#include <omp.h>

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

/* ------------------------------------------------------------------ */
/* our fp32 log2 definition (DESIGN.md "discrete decisions")           */
/* The reference's Log2 is GLSL.std.450 (implementation-defined         */
/* precision) on the GPU; the cluster slice is a truncation of it, so   */
/* oracle and kernel share one exactly specified evaluation:            */
/*   x = m * 2^e, m in (sqrt(1/2), sqrt(2)];  s = (m-1)/(m+1)           */
/*   ln m = 2s(1 + z/3 + z^2/5 + z^3/7 + z^4/9 + z^5/11), z = s^2       */
/* evaluated in fp32, Horner, no FMA.                                   */
/* ------------------------------------------------------------------ */
float orc_log2_spec(float x) {
    uint32_t bits;
    memcpy(&bits, &x, 4);
    if (!(x > 0.0f) || bits >= 0x7f800000u) {
        if (x == 0.0f) return -INFINITY;
        if (x > 0.0f) return x; /* +inf */
        return NAN;
    }
    int e = 0;
    if (bits < 0x00800000u) { /* subnormal: scale by 2^24 */
        x = x * 16777216.0f;
        memcpy(&bits, &x, 4);
        e = -24;
    }
    e += (int)(bits >> 23) - 127;
    uint32_t mb = (bits & 0x007fffffu) | 0x3f800000u;
    float m;
    memcpy(&m, &mb, 4);
    if (m > 1.41421354f) {
        m = m * 0.5f;
        e = e + 1;
    }
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float p = z * 0.0909090936f + 0.111111112f;
    p = z * p + 0.142857149f;
    p = z * p + 0.2f;
    p = z * p + 0.333333343f;
    float s2 = s + s;
    float r = s2 + s2 * (z * p);
    return (float)e + r * 1.44269502f;
}

/* shared-structs/src/lib.rs:44-52 */
void orc_light_cluster_coefficients_new(float z_near, float z_far, uint32_t slices, tr_light_cluster_coefficients* out) {
    out->z_near = z_near;
    out->z_far = z_far;
    out->num_depth_slices = slices;
    out->scale = (float)slices / log2f(z_far / z_near);
    out->bias = -((float)slices * log2f(z_near) / log2f(z_far / z_near));
}

/* shared-structs/src/lib.rs:54-58 */
float orc_linear_depth(const tr_light_cluster_coefficients* c, float frag_depth) {
    float depth_range = 2.0f * (1.0f - frag_depth) - 1.0f;
    return 2.0f * c->z_near * c->z_far / (c->z_far + c->z_near - depth_range * (c->z_far - c->z_near));
}

/* Rust `f32 as u32`: saturating, NaN -> 0 */
static uint32_t f32_as_u32(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)v;
}

/* shared-structs/src/lib.rs:61-63 */
uint32_t orc_get_depth_slice(const tr_light_cluster_coefficients* c, float frag_depth) {
    float v = orc_log2_spec(orc_linear_depth(c, frag_depth)) * c->scale + c->bias;
    return f32_as_u32(f_max(v, 0.0f));
}

/* shared-structs/src/lib.rs:129-138 */
float orc_spotlight_factor(const tr_light* l, v3 direction_to_light) {
    v3 spot_dir = v3_new(l->spotlight_direction_and_outer_angle.x, l->spotlight_direction_and_outer_angle.y,
                         l->spotlight_direction_and_outer_angle.z);
    float theta = v3_dot(v3_neg(direction_to_light), spot_dir);
    float outer_angle = l->spotlight_direction_and_outer_angle.w;
    float epsilon = l->position_and_spotlight_epsilon.w;
    float cos_outer = (float)cos((double)outer_angle);
    return f_max((theta - cos_outer) / epsilon, 0.0f);
}

/* ------------------------------------------------------------------ */
/* Samplers (sampler state src/main.rs:683-705)   */
/* ------------------------------------------------------------------ */
static int64_t clamp_i64(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* texel-space setup shared by both samplers: p = u*size - 0.5 */
static void bilinear_setup(float u, uint32_t size, uint32_t* i0, uint32_t* i1, float* frac) {
    float p = u * (float)size - 0.5f;
    if (!(p == p)) p = 0.0f; /* NaN coordinate -> texel 0 */
    float fl = floorf(p);
    float f = p - fl;
    int64_t i;
    if (fl <= -2147483648.0f) {
        i = -2147483648LL;
        f = 0.0f;
    } else if (fl >= 2147483520.0f) {
        i = 2147483520LL;
        f = 0.0f;
    } else {
        i = (int64_t)fl;
    }
    *i0 = (uint32_t)clamp_i64(i, 0, (int64_t)size - 1);
    *i1 = (uint32_t)clamp_i64(i + 1, 0, (int64_t)size - 1);
    *frac = f;
}

static float lerpf(float a, float b, float t) { return a + (b - a) * t; }

static v3 sample_level(const orc_pyramid* p, uint32_t level, float u, float v) {
    uint32_t w = p->width[level], h = p->height[level];
    const uint16_t* d = p->data[level];
    uint32_t x0, x1, y0, y1;
    float fx, fy;
    bilinear_setup(u, w, &x0, &x1, &fx);
    bilinear_setup(v, h, &y0, &y1, &fy);
    float out[3];
    for (int c = 0; c < 3; c++) {
        float t00 = f16_bits_to_f32(d[((size_t)y0 * w + x0) * 4 + c]);
        float t10 = f16_bits_to_f32(d[((size_t)y0 * w + x1) * 4 + c]);
        float t01 = f16_bits_to_f32(d[((size_t)y1 * w + x0) * 4 + c]);
        float t11 = f16_bits_to_f32(d[((size_t)y1 * w + x1) * 4 + c]);
        float top = lerpf(t00, t10, fx);
        float bot = lerpf(t01, t11, fx);
        out[c] = lerpf(top, bot, fy);
    }
    return v3_new(out[0], out[1], out[2]);
}

/* framebuffer.sample_by_lod(clamp_sampler, uv, lod).rgb — shader/src/lib.rs:135-138 */
v3 orc_sample_pyramid(const orc_pyramid* p, float u, float v, float lod) {
    float max_lod = (float)(p->levels - 1);
    if (!(lod > 0.0f)) lod = 0.0f;
    if (lod > max_lod) lod = max_lod;
    float l0f = floorf(lod);
    uint32_t l0 = (uint32_t)l0f;
    uint32_t l1 = l0 + 1 < p->levels ? l0 + 1 : p->levels - 1;
    float t = lod - l0f;
    v3 s0 = sample_level(p, l0, u, v);
    v3 s1 = sample_level(p, l1, u, v);
    return v3_new(lerpf(s0.x, s1.x, t), lerpf(s0.y, s1.y, t), lerpf(s0.z, s1.z, t));
}

/* textures[ggx_lut].sample(clamp_sampler, (n.v, roughness)).xy — shader/src/lib.rs:126-133;
 * RGBA8 UNORM, one level (src/main.rs:295-330) */
v2 orc_sample_lut(const orc_lut* lut, float n_dot_v, float roughness) {
    uint32_t x0, x1, y0, y1;
    float fx, fy;
    bilinear_setup(n_dot_v, lut->width, &x0, &x1, &fx);
    bilinear_setup(roughness, lut->height, &y0, &y1, &fy);
    float out[2];
    for (int c = 0; c < 2; c++) {
        float t00 = (float)lut->rgba8[((size_t)y0 * lut->width + x0) * 4 + c] / 255.0f;
        float t10 = (float)lut->rgba8[((size_t)y0 * lut->width + x1) * 4 + c] / 255.0f;
        float t01 = (float)lut->rgba8[((size_t)y1 * lut->width + x0) * 4 + c] / 255.0f;
        float t11 = (float)lut->rgba8[((size_t)y1 * lut->width + x1) * 4 + c] / 255.0f;
        out[c] = lerpf(lerpf(t00, t10, fx), lerpf(t01, t11, fx), fy);
    }
    v2 r = {out[0], out[1]};
    return r;
}

/* ------------------------------------------------------------------ */
/* material textures: texture.sample(sampler, uv), shader/src/lib.rs:251-267 */
/* sampler = linear min/mag, linear mip, REPEAT (default address mode),  */
/* min_lod 0, max_lod NONE, no anisotropy (src/main.rs:683-692)          */
/* ------------------------------------------------------------------ */
float orc_srgb8_to_linear(uint8_t c) { /* R8G8B8A8_SRGB decode, evaluated in double and rounded once */
    double x = (double)c / 255.0;
    return (float)(x <= 0.04045 ? x / 12.92 : pow((x + 0.055) / 1.055, 2.4));
}

static void repeat_setup(float u, uint32_t size, uint32_t* i0, uint32_t* i1, float* frac) {
    float p = u * (float)size - 0.5f;
    if (!(p == p) || !(fabsf(p) < 1.0e9f)) p = 0.0f; /* NaN / absurd coordinate -> texel 0 */
    float fl = floorf(p);
    int64_t i = (int64_t)fl, n = (int64_t)size;
    int64_t a = i % n, b = (i + 1) % n;
    *i0 = (uint32_t)(a < 0 ? a + n : a);
    *i1 = (uint32_t)(b < 0 ? b + n : b);
    *frac = p - fl;
}

static v4 sample_texture_level(const orc_texture* t, uint32_t level, v2 uv) {
    uint32_t w = t->width >> level, h = t->height >> level;
    if (w == 0) w = 1;
    if (h == 0) h = 1;
    const uint8_t* d = t->data[level];
    uint32_t x0, x1, y0, y1;
    float fx, fy;
    repeat_setup(uv.x, w, &x0, &x1, &fx);
    repeat_setup(uv.y, h, &y0, &y1, &fy);
    float out[4];
    for (int c = 0; c < 4; c++) {
        uint8_t b00 = d[((size_t)y0 * w + x0) * 4 + c], b10 = d[((size_t)y0 * w + x1) * 4 + c];
        uint8_t b01 = d[((size_t)y1 * w + x0) * 4 + c], b11 = d[((size_t)y1 * w + x1) * 4 + c];
        int decode = t->srgb && c < 3; /* alpha is linear in the sRGB formats */
        float t00 = decode ? orc_srgb8_to_linear(b00) : (float)b00 / 255.0f;
        float t10 = decode ? orc_srgb8_to_linear(b10) : (float)b10 / 255.0f;
        float t01 = decode ? orc_srgb8_to_linear(b01) : (float)b01 / 255.0f;
        float t11 = decode ? orc_srgb8_to_linear(b11) : (float)b11 / 255.0f;
        out[c] = lerpf(lerpf(t00, t10, fx), lerpf(t01, t11, fx), fy);
    }
    return v4_new(out[0], out[1], out[2], out[3]);
}

v4 orc_sample_texture(const orc_texture* t, v2 uv, v2 duv_dx, v2 duv_dy) {
    /* scale factor rho of the Vulkan level-of-detail operation, in level-0 texels per pixel */
    float ux = duv_dx.x * (float)t->width, vx = duv_dx.y * (float)t->height;
    float uy = duv_dy.x * (float)t->width, vy = duv_dy.y * (float)t->height;
    float rho2 = f_max(ux * ux + vx * vx, uy * uy + vy * vy);
    float lod = 0.5f * orc_log2_spec(rho2); /* log2(rho); -inf for rho == 0 */
    float max_lod = (float)(t->levels - 1);
    if (!(lod > 0.0f)) lod = 0.0f;
    if (lod > max_lod) lod = max_lod;
    float l0f = floorf(lod);
    uint32_t l0 = (uint32_t)l0f;
    uint32_t l1 = l0 + 1 < t->levels ? l0 + 1 : t->levels - 1;
    float f = lod - l0f;
    v4 a = sample_texture_level(t, l0, uv), b = sample_texture_level(t, l1, uv);
    return v4_new(lerpf(a.x, b.x, f), lerpf(a.y, b.y, f), lerpf(a.z, b.z, f), lerpf(a.w, b.w, f));
}

/* ------------------------------------------------------------------ */
/* fragment stage                                                       */
/* ------------------------------------------------------------------ */
static v3 from_a(tr_vec3a a) { return v3_new(a.x, a.y, a.z); }

typedef struct { /* TextureSampler, shader/src/lib.rs:251-267 */
    const orc_scene* s;
    v2 uv, duv_dx, duv_dy;
} texture_sampler;

static v4 tex_sample(const texture_sampler* ts, int32_t id) {
    if (!ts->s->textures || (uint32_t)id >= ts->s->n_textures) return v4_new(0.0f, 0.0f, 0.0f, 0.0f); /* robust access */
    return orc_sample_texture(&ts->s->textures[id], ts->uv, ts->duv_dx, ts->duv_dy);
}

/* lighting.rs:261-301 */
static orc_material_params get_material_params(tr_vec4 diffuse, const tr_material_info* m, const texture_sampler* ts) {
    orc_material_params r;
    float metallic = m->metallic_factor, roughness = m->roughness_factor;
    if (m->textures.metallic_roughness != -1) {
        v4 sample = tex_sample(ts, m->textures.metallic_roughness);
        metallic *= sample.z; /* "These two are switched!" lighting.rs:272-276 */
        roughness *= sample.y;
    }
    v3 specular_colour = from_a(m->specular_colour_factor);
    if (m->textures.specular_colour != -1) {
        v4 sample = tex_sample(ts, m->textures.specular_colour);
        specular_colour = v3_mul(specular_colour, v3_new(sample.x, sample.y, sample.z));
    }
    float specular_factor = m->specular_factor;
    if (m->textures.specular != -1) specular_factor *= tex_sample(ts, m->textures.specular).w;
    r.diffuse_colour = v3_new(diffuse.x, diffuse.y, diffuse.z);
    r.metallic = metallic;
    r.perceptual_roughness = roughness;
    r.index_of_refraction = m->index_of_refraction;
    r.specular_colour = specular_colour;
    r.specular_factor = specular_factor;
    return r;
}

/* lighting.rs:303-313 */
static v3 get_emission(const tr_material_info* m, const texture_sampler* ts) {
    v3 e = from_a(m->emissive_factor);
    if (m->textures.emissive != -1) {
        v4 sample = tex_sample(ts, m->textures.emissive);
        e = v3_mul(e, v3_new(sample.x, sample.y, sample.z));
    }
    return e;
}

/* lighting.rs:222-259: calculate_normal + compute_cotangent_frame; `position` there is -view_vector, whose ddx / ddy are
 * the world-position differences */
static v3 calculate_normal(v3 interpolated_normal, const texture_sampler* ts, const tr_material_info* m,
                           const orc_frag_derivatives* d) {
    v3 normal = v3_normalize(interpolated_normal);
    if (m->textures.normal_map != -1) {
        v4 smp = tex_sample(ts, m->textures.normal_map);
        float k = 128.0f / 127.0f;
        v3 map_normal = v3_new(smp.x * 255.0f / 127.0f - k, smp.y * 255.0f / 127.0f - k, smp.z * 255.0f / 127.0f - k);
        v3 dp1 = d->dpos_dx, dp2 = d->dpos_dy;
        v3 dp2perp = v3_cross(dp2, normal), dp1perp = v3_cross(normal, dp1);
        v3 t = v3_add(v3_scale(dp2perp, d->duv_dx.x), v3_scale(dp1perp, d->duv_dy.x));
        v3 b = v3_add(v3_scale(dp2perp, d->duv_dx.y), v3_scale(dp1perp, d->duv_dy.y));
        float invmax = 1.0f / sqrtf(f_max(v3_dot(t, t), v3_dot(b, b)));
        v3 c0 = v3_scale(t, invmax), c1 = v3_scale(b, invmax);
        v3 r = v3_add(v3_add(v3_scale(c0, map_normal.x), v3_scale(c1, map_normal.y)), v3_scale(normal, map_normal.z));
        normal = v3_normalize(r);
    }
    return normal;
}

static const orc_frag_derivatives k_zero_derivatives = {{0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};

/* shader/src/lib.rs:88-98 / 205-215; robust-buffer-access semantics for an out-of-range cluster */
static uint32_t cluster_index(v4 frag_coord, const tr_uniforms* u) {
    uint32_t cx = f32_as_u32(frag_coord.x / u->cluster_size_in_pixels.x);
    uint32_t cy = f32_as_u32(frag_coord.y / u->cluster_size_in_pixels.y);
    uint32_t cz = orc_get_depth_slice(&u->light_clustering_coefficients, frag_coord.z);
    return cz * u->num_clusters.x * u->num_clusters.y + cy * u->num_clusters.x + cx;
}

static v3 light_position(const tr_light* l) {
    return v3_new(l->position_and_spotlight_epsilon.x, l->position_and_spotlight_epsilon.y,
                  l->position_and_spotlight_epsilon.z);
}
static v3 light_emission(const tr_light* l) {
    return v3_new(l->colour_emission_and_falloff_distance_sq.x, l->colour_emission_and_falloff_distance_sq.y,
                  l->colour_emission_and_falloff_distance_sq.z);
}

/* shader/src/lib.rs:164-249 + lighting.rs:145-220 */
v4 orc_fragment(v3 position, v3 normal_in, v2 uv, uint32_t material_id, v4 frag_coord, const orc_scene* s,
                const orc_frag_derivatives* d) {
    if (!d) d = &k_zero_derivatives;
    const tr_material_info* material = &s->materials[material_id];
    texture_sampler ts = {s, uv, d->duv_dx, d->duv_dy};
    tr_vec4 diffuse = material->diffuse_factor;
    if (material->textures.diffuse != -1) { /* lib.rs:190-194 */
        v4 smp = tex_sample(&ts, material->textures.diffuse);
        diffuse.x *= smp.x; diffuse.y *= smp.y; diffuse.z *= smp.z; diffuse.w *= smp.w;
    }

    v3 view_vector = v3_sub(from_a(s->pc->view_position), position);
    v3 view = v3_normalize(view_vector);
    v3 normal = calculate_normal(normal_in, &ts, material, d); /* lighting.rs:222-241 */
    orc_material_params mp = get_material_params(diffuse, material, &ts);
    v3 emission = get_emission(material, &ts); /* lighting.rs:303-313 */

    uint32_t cluster = cluster_index(frag_coord, s->uniforms);
    uint32_t num_lights = cluster < s->n_clusters ? s->cluster_light_counts[cluster] : 0u;

    v3 sun_dir = from_a(s->uniforms->sun_dir);
    float sun_factor = 1.0f;
    if (s->accel) /* lighting.rs:154-165; "todo: ambient lighting via probes or idk!" */
        sun_factor = f_max(orc_trace_shadow(s->accel, position, sun_dir, 10000.0f), 0.1f);
    v3 sun_intensity = v3_scale(from_a(s->uniforms->sun_intensity), sun_factor); /* lighting.rs:168-173 */
    orc_brdf_result sum = orc_basic_brdf(normal, sun_dir, sun_intensity, view, mp);

    uint32_t offset = cluster * TR_MAX_LIGHTS_PER_CLUSTER;
    for (uint32_t i = 0; i < num_lights; i++) {
        const tr_light* light = &s->lights[s->cluster_light_indices[offset + i]];
        v3 direction;
        float distance, attenuation;
        orc_light_direction_and_attenuation(position, light_position(light), &direction, &distance, &attenuation);
        float factor = 1.0f;
        if (s->accel) factor *= orc_trace_shadow(s->accel, position, direction, distance); /* lighting.rs:186-195 */
        if (light->spotlight_direction_and_outer_angle.w != 0.0f) factor *= orc_spotlight_factor(light, direction);
        v3 emission_l = v3_scale(light_emission(light), factor);
        orc_brdf_result r = orc_basic_brdf(normal, direction, v3_scale(emission_l, attenuation), view, mp);
        sum.diffuse = v3_add(sum.diffuse, r.diffuse);
        sum.specular = v3_add(sum.specular, r.specular);
    }

    v3 out = v3_add(v3_add(sum.diffuse, sum.specular), emission);
    return v4_new(out.x, out.y, out.z, 1.0f);
}

/* shader/src/lib.rs:37-162 + lighting.rs:13-95 */
v4 orc_fragment_transmission(v3 position, v3 normal_in, v2 uv, uint32_t material_id, float model_scale, v4 frag_coord,
                             const orc_scene* s, const orc_pyramid* fb, const orc_lut* lut, const orc_frag_derivatives* d) {
    if (!d) d = &k_zero_derivatives;
    const tr_material_info* material = &s->materials[material_id];
    texture_sampler ts = {s, uv, d->duv_dx, d->duv_dy};
    tr_vec4 diffuse = material->diffuse_factor;
    if (material->textures.diffuse != -1) { /* lib.rs:66-69 */
        v4 smp = tex_sample(&ts, material->textures.diffuse);
        diffuse.x *= smp.x; diffuse.y *= smp.y; diffuse.z *= smp.z; diffuse.w *= smp.w;
    }
    float transmission_factor = material->transmission_factor;
    if (material->textures.transmission != -1) transmission_factor *= tex_sample(&ts, material->textures.transmission).x; /* :71-77 */

    v3 view_vector = v3_sub(from_a(s->pc->view_position), position);
    v3 view = v3_normalize(view_vector);
    v3 normal = calculate_normal(normal_in, &ts, material, d);
    orc_material_params mp = get_material_params(diffuse, material, &ts);
    v3 emission = get_emission(material, &ts);

    uint32_t cluster = cluster_index(frag_coord, s->uniforms);
    uint32_t num_lights = cluster < s->n_clusters ? s->cluster_light_counts[cluster] : 0u;

    /* lighting.rs:37-53 */
    v3 sun_dir = from_a(s->uniforms->sun_dir);
    float sun_factor = s->accel ? orc_trace_shadow(s->accel, position, sun_dir, 10000.0f) : 1.0f; /* lighting.rs:25-35 */
    v3 sun_intensity = v3_scale(from_a(s->uniforms->sun_intensity), sun_factor);
    orc_brdf_result sum = orc_basic_brdf(normal, sun_dir, sun_intensity, view, mp);
    v3 transmission = v3_mul(sun_intensity, orc_transmission_btdf(mp, normal, view, sun_dir));

    /* lighting.rs:58-92 — note: no spotlight factor in this loop */
    uint32_t offset = cluster * TR_MAX_LIGHTS_PER_CLUSTER;
    for (uint32_t i = 0; i < num_lights; i++) {
        const tr_light* light = &s->lights[s->cluster_light_indices[offset + i]];
        v3 direction;
        float distance, attenuation;
        orc_light_direction_and_attenuation(position, light_position(light), &direction, &distance, &attenuation);
        float factor = s->accel ? orc_trace_shadow(s->accel, position, direction, distance) : 1.0f; /* lighting.rs:64-75 */
        v3 emission_l = v3_scale(light_emission(light), factor);
        orc_brdf_result r = orc_basic_brdf(normal, direction, v3_scale(emission_l, attenuation), view, mp);
        sum.diffuse = v3_add(sum.diffuse, r.diffuse);
        sum.specular = v3_add(sum.specular, r.specular);
        transmission = v3_add(transmission, v3_mul(v3_scale(emission_l, attenuation),
                                                   orc_transmission_btdf(mp, normal, view, direction)));
    }

    float thickness = material->thickness_factor; /* lib.rs:120-124 */
    if (material->textures.thickness != -1) thickness *= tex_sample(&ts, material->textures.thickness).y;

    orc_ibl_params ip;
    ip.material_params = mp;
    ip.framebuffer_size_x = s->pc->framebuffer_size.x;
    ip.normal = normal;
    ip.view = view;
    memcpy(&ip.proj_view_matrix, &s->pc->proj_view, sizeof(m4));
    ip.position = position;
    ip.thickness = thickness;
    ip.model_scale = model_scale;
    ip.attenuation_distance = material->attenuation_distance;
    ip.attenuation_colour = from_a(material->attenuation_colour);
    transmission = v3_add(transmission, orc_ibl_volume_refraction(&ip, fb, lut)); /* lib.rs:140-155 */

    v3 real_transmission = v3_scale(transmission, transmission_factor);          /* lib.rs:157 */
    v3 diffuse_out = v3_lerp(sum.diffuse, real_transmission, transmission_factor); /* lib.rs:159 */
    v3 out = v3_add(v3_add(diffuse_out, sum.specular), emission);                 /* lib.rs:161 */
    return v4_new(out.x, out.y, out.z, 1.0f);
}

/* ------------------------------------------------------------------ */
/* G-buffer decode: what the rasteriser + varying interpolation feed    */
/* the fragment stage (ours to define; DESIGN.md "G-buffer")            */
/* ------------------------------------------------------------------ */
static v3 decode_position(const orc_gbuffer* g, const m4* inv_pv, uint32_t x, uint32_t y, float depth) {
    size_t i = (size_t)y * g->width + x;
    if (g->position) return v3_new(g->position[i * 3], g->position[i * 3 + 1], g->position[i * 3 + 2]);
    float ndc_x = ((float)x + 0.5f) / (float)g->width * 2.0f - 1.0f;
    float ndc_y = ((float)y + 0.5f) / (float)g->height * 2.0f - 1.0f;
    v4 h = m4_mul_v4(inv_pv, v4_new(ndc_x, ndc_y, depth, 1.0f));
    return v3_new(h.x / h.w, h.y / h.w, h.z / h.w);
}

/* differences to the right / lower neighbour on the pixel's own triangle, from the derivative planes of the G-buffer */
static orc_frag_derivatives decode_derivatives(const orc_gbuffer* g, const m4* inv_pv, uint32_t x, uint32_t y, float depth, v3 pos) {
    orc_frag_derivatives d = k_zero_derivatives;
    size_t i = (size_t)y * g->width + x;
    if (g->duv) {
        d.duv_dx.x = g->duv[i * 4]; d.duv_dx.y = g->duv[i * 4 + 1];
        d.duv_dy.x = g->duv[i * 4 + 2]; d.duv_dy.y = g->duv[i * 4 + 3];
    }
    if (g->ddepth && !g->position) {
        v3 px = decode_position(g, inv_pv, x + 1, y, depth + g->ddepth[i * 2]);
        v3 py = decode_position(g, inv_pv, x, y + 1, depth + g->ddepth[i * 2 + 1]);
        d.dpos_dx = v3_sub(px, pos);
        d.dpos_dy = v3_sub(py, pos);
    }
    return d;
}

static void store_px(float* f32buf, uint16_t* a, uint16_t* b, size_t i, v4 c) {
    if (f32buf) {
        f32buf[i * 4] = c.x; f32buf[i * 4 + 1] = c.y; f32buf[i * 4 + 2] = c.z; f32buf[i * 4 + 3] = c.w;
    }
    uint16_t h[4] = {f32_to_f16_bits(c.x), f32_to_f16_bits(c.y), f32_to_f16_bits(c.z), f32_to_f16_bits(c.w)};
    if (a) memcpy(a + i * 4, h, 8);
    if (b) memcpy(b + i * 4, h, 8);
}

void orc_shade_opaque_frame(const orc_gbuffer* g, const orc_scene* s, uint32_t y0, uint32_t y1, float* hdr_f32,
                            uint16_t* hdr_f16, uint16_t* opaque_f16) {
    tr_mat4 inv;
    orc_mat4_inverse(&s->pc->proj_view, &inv);
    m4 inv_pv;
    memcpy(&inv_pv, &inv, sizeof(m4));
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
        uint32_t y = (uint32_t)yy;
        for (uint32_t x = 0; x < g->width; x++) {
            size_t i = (size_t)y * g->width + x;
            float depth = g->depth[i];
            v4 c = v4_new(0.0f, 0.0f, 0.0f, 1.0f); /* clear colour, main.rs:1592-1602 */
            if (depth != 0.0f) {
                v3 pos = decode_position(g, &inv_pv, x, y, depth);
                v3 n = v3_new(g->normal[i * 3], g->normal[i * 3 + 1], g->normal[i * 3 + 2]);
                v2 uv = {g->uv ? g->uv[i * 2] : 0.0f, g->uv ? g->uv[i * 2 + 1] : 0.0f};
                v4 fc = v4_new((float)x + 0.5f, (float)y + 0.5f, depth, 1.0f);
                orc_frag_derivatives d = decode_derivatives(g, &inv_pv, x, y, depth, pos);
                c = orc_fragment(pos, n, uv, g->material_id[i], fc, s, &d);
            }
            store_px(hdr_f32, hdr_f16, opaque_f16, i, c);
        }
    }
}

void orc_shade_transmission_frame(const orc_gbuffer* g, const orc_scene* s, const orc_pyramid* fb, const orc_lut* lut,
                                  uint32_t y0, uint32_t y1, float* hdr_f32, uint16_t* hdr_f16) {
    tr_mat4 inv;
    orc_mat4_inverse(&s->pc->proj_view, &inv);
    m4 inv_pv;
    memcpy(&inv_pv, &inv, sizeof(m4));
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
        uint32_t y = (uint32_t)yy;
        for (uint32_t x = 0; x < g->width; x++) {
            size_t i = (size_t)y * g->width + x;
            float depth = g->depth[i];
            if (depth == 0.0f) continue; /* render pass LOADs hdr; uncovered pixels keep the opaque result */
            v3 pos = decode_position(g, &inv_pv, x, y, depth);
            v3 n = v3_new(g->normal[i * 3], g->normal[i * 3 + 1], g->normal[i * 3 + 2]);
            v2 uv = {g->uv ? g->uv[i * 2] : 0.0f, g->uv ? g->uv[i * 2 + 1] : 0.0f};
            v4 fc = v4_new((float)x + 0.5f, (float)y + 0.5f, depth, 1.0f);
            float scale = g->scale ? g->scale[i] : 1.0f;
            orc_frag_derivatives d = decode_derivatives(g, &inv_pv, x, y, depth, pos);
            v4 c = orc_fragment_transmission(pos, n, uv, g->material_id[i], scale, fc, s, fb, lut, &d);
            store_px(hdr_f32, hdr_f16, NULL, i, c);
        }
    }
}

void orc_shade_frame_with(const orc_gbuffer* g, const tr_push_constants* pc, int transmissive, uint32_t y0, uint32_t y1,
                          orc_fragment_fn fn, void* user, float* hdr_f32, uint16_t* hdr_f16, uint16_t* opaque_f16) {
    tr_mat4 inv;
    orc_mat4_inverse(&pc->proj_view, &inv);
    m4 inv_pv;
    memcpy(&inv_pv, &inv, sizeof(m4));
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
        uint32_t y = (uint32_t)yy;
        for (uint32_t x = 0; x < g->width; x++) {
            size_t i = (size_t)y * g->width + x;
            float depth = g->depth[i];
            v4 c = v4_new(0.0f, 0.0f, 0.0f, 1.0f);
            if (depth == 0.0f && transmissive) continue;
            if (depth != 0.0f) {
                v3 pos = decode_position(g, &inv_pv, x, y, depth);
                v3 n = v3_new(g->normal[i * 3], g->normal[i * 3 + 1], g->normal[i * 3 + 2]);
                v2 uv = {g->uv ? g->uv[i * 2] : 0.0f, g->uv ? g->uv[i * 2 + 1] : 0.0f};
                v4 fc = v4_new((float)x + 0.5f, (float)y + 0.5f, depth, 1.0f);
                float scale = g->scale ? g->scale[i] : 1.0f;
                orc_frag_derivatives d = decode_derivatives(g, &inv_pv, x, y, depth, pos);
                c = fn(pos, n, uv, g->material_id[i], scale, fc, &d, user);
            }
            store_px(hdr_f32, hdr_f16, transmissive ? NULL : opaque_f16, i, c);
        }
    }
}

/* Which shadow rays of a layer are occluded, as the product's shadow pass stores them: plane 0..3 = bit i set when the
 * ray towards the i-th light of the pixel's cluster list is occluded, plane 4 bit 0 = the sun ray.  mask: [5][h*w]. */
void orc_shadow_mask_frame(const orc_gbuffer* g, const orc_scene* s, uint32_t y0, uint32_t y1, uint32_t* mask) {
    tr_mat4 inv;
    orc_mat4_inverse(&s->pc->proj_view, &inv);
    m4 inv_pv;
    memcpy(&inv_pv, &inv, sizeof(m4));
    const size_t plane = (size_t)g->width * g->height;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
        uint32_t y = (uint32_t)yy;
        for (uint32_t x = 0; x < g->width; x++) {
            size_t i = (size_t)y * g->width + x;
            float depth = g->depth[i];
            uint32_t m[5] = {0, 0, 0, 0, 0};
            if (depth != 0.0f && s->accel) {
                v3 pos = decode_position(g, &inv_pv, x, y, depth);
                v4 fc = v4_new((float)x + 0.5f, (float)y + 0.5f, depth, 1.0f);
                uint32_t cluster = cluster_index(fc, s->uniforms);
                uint32_t num_lights = cluster < s->n_clusters ? s->cluster_light_counts[cluster] : 0u;
                if (orc_trace_shadow(s->accel, pos, from_a(s->uniforms->sun_dir), 10000.0f) == 0.0f) m[4] = 1u;
                for (uint32_t k = 0; k < num_lights; k++) {
                    const tr_light* light = &s->lights[s->cluster_light_indices[cluster * TR_MAX_LIGHTS_PER_CLUSTER + k]];
                    v3 direction;
                    float distance, attenuation;
                    orc_light_direction_and_attenuation(pos, light_position(light), &direction, &distance, &attenuation);
                    if (orc_trace_shadow(s->accel, pos, direction, distance) == 0.0f) m[k >> 5] |= 1u << (k & 31u);
                }
            }
            for (int k = 0; k < 5; k++) mask[(size_t)k * plane + i] = m[k];
        }
    }
}

/* oracle/pbr.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of /root/reference/glam-pbr/src/lib.rs (whole file) in plain C.
 * Every function cites the lines it follows; operation order follows the Rust
 * expression trees (left-to-right, operator precedence) so fp32 rounding is
 * the reference's to within libm differences.  PARITY: pinned through shade.c's callers to the reference's compiled
 * fragment.spv / fragment_transmission.spv, bit-equal fp32 pixels (oracle.h, tests/test_reference_spirv.py).
 */
#include "oracle.h"

#include <float.h>

#define ORC_PI 3.14159265358979323846264338327950288f        /* core::f32::consts::PI */
#define ORC_FRAC_1_PI 0.318309886183790671537767526745028724f /* core::f32::consts::FRAC_1_PI */

/* glam-pbr/src/lib.rs:25-27 */
static float clampf(float value, float lo, float hi) { return f_min(f_max(value, lo), hi); }

/* glam-pbr/src/lib.rs:92-99 — Dot::new clamps to f32::EPSILON */
static float shading_dot(v3 a, v3 b) { return f_max(v3_dot(a, b), FLT_EPSILON); }

/* glam-pbr/src/lib.rs:12-23 */
void orc_light_direction_and_attenuation(v3 fragment_position, v3 light_position, v3* direction, float* distance,
                                         float* attenuation) {
    v3 vector = v3_sub(light_position, fragment_position);
    float distance_sq = v3_length_squared(vector);
    float dist = sqrtf(distance_sq);
    *direction = v3_divs(vector, dist);
    *distance = dist;
    *attenuation = 1.0f / distance_sq;
}

/* glam-pbr/src/lib.rs:101-109 */
float orc_d_ggx(float noh, float alpha) {
    float alpha_roughness_sq = alpha * alpha;
    float f = (noh * noh) * (alpha_roughness_sq - 1.0f) + 1.0f;
    return alpha_roughness_sq / (ORC_PI * f * f);
}

/* glam-pbr/src/lib.rs:114-133 */
float orc_v_smith_ggx_correlated(float nov, float nol, float alpha) {
    float a2 = alpha * alpha;
    float ggx_v = nol * sqrtf(nov * nov * (1.0f - a2) + a2);
    float ggx_l = nov * sqrtf(nol * nol * (1.0f - a2) + a2);
    float ggx = ggx_v + ggx_l;
    if (ggx > 0.0f) return 0.5f / ggx;
    return 0.0f;
}

/* glam-pbr/src/lib.rs:137-139 */
v3 orc_fresnel_schlick(float vdoth, v3 f0, v3 f90) {
    float p = powf(1.0f - vdoth, 5.0f);
    return v3_add(f0, v3_scale(v3_sub(f90, f0), p));
}

/* glam-pbr/src/lib.rs:192-195 */
float orc_ior_to_dielectric_f0(float ior) {
    float root = (ior - 1.0f) / (ior + 1.0f);
    return root * root;
}

/* glam-pbr/src/lib.rs:425-430 */
v3 orc_calculate_combined_f0(orc_material_params m) {
    v3 dielectric =
        v3_scale(v3_scale(m.specular_colour, orc_ior_to_dielectric_f0(m.index_of_refraction)), m.specular_factor);
    return v3_lerp(dielectric, m.diffuse_colour, m.metallic);
}

/* glam-pbr/src/lib.rs:432-435 */
v3 orc_calculate_combined_f90(orc_material_params m) {
    return v3_lerp(v3_splat(m.specular_factor), v3_splat(1.0f), m.metallic);
}

/* glam-pbr/src/lib.rs:356-360 */
static v3 diffuse_brdf(v3 base, v3 fresnel) {
    return v3_scale(base, (1.0f - v3_max_element(fresnel)) * ORC_FRAC_1_PI);
}

/* glam-pbr/src/lib.rs:362-375 */
static v3 specular_brdf(float nov, float nol, float noh, float alpha, v3 fresnel) {
    float d = orc_d_ggx(noh, alpha);
    float g = orc_v_smith_ggx_correlated(nov, nol, alpha);
    return v3_scale(fresnel, d * g);
}

/* glam-pbr/src/lib.rs:377-423 */
orc_brdf_result orc_basic_brdf(v3 normal, v3 light, v3 light_intensity, v3 view, orc_material_params m) {
    float actual_roughness = m.perceptual_roughness * m.perceptual_roughness; /* :154-156 */
    v3 halfway = v3_normalize(v3_add(view, light));                            /* :64-68 */
    float noh = shading_dot(normal, halfway);
    float nov = shading_dot(normal, view);
    float nol = shading_dot(normal, light);
    float voh = shading_dot(view, halfway);

    v3 c_diff = v3_lerp(m.diffuse_colour, v3_splat(0.0f), m.metallic);
    v3 f0 = orc_calculate_combined_f0(m);
    v3 f90 = orc_calculate_combined_f90(m);
    v3 fresnel = orc_fresnel_schlick(voh, f0, f90);

    orc_brdf_result r;
    r.diffuse = v3_mul(v3_scale(light_intensity, nol), diffuse_brdf(c_diff, fresnel));
    r.specular = v3_mul(v3_scale(light_intensity, nol), specular_brdf(nov, nol, noh, actual_roughness, fresnel));
    return r;
}

/* glam-pbr/src/lib.rs:200-233 */
v3 orc_transmission_btdf(orc_material_params m, v3 normal, v3 view, v3 light) {
    float actual_roughness = m.perceptual_roughness * m.perceptual_roughness;
    /* :144-148 ActualRoughness::apply_ior */
    float transmission_roughness = actual_roughness * clampf(m.index_of_refraction * 2.0f - 2.0f, 0.0f, 1.0f);

    /* :211  light + 2.0 * normal * dot(-light, normal) */
    v3 mirrored =
        v3_normalize(v3_add(light, v3_scale(v3_scale(normal, 2.0f), v3_dot(v3_neg(light), normal))));

    v3 halfway = v3_normalize(v3_add(view, mirrored));
    float noh = shading_dot(normal, halfway);
    float voh = shading_dot(view, halfway);
    float nov = shading_dot(normal, view);
    float nolm = shading_dot(normal, mirrored);

    float distribution = orc_d_ggx(noh, transmission_roughness);
    float shadowing = orc_v_smith_ggx_correlated(nov, nolm, transmission_roughness);

    v3 f0 = orc_calculate_combined_f0(m);
    v3 f90 = orc_calculate_combined_f90(m);
    v3 fresnel = orc_fresnel_schlick(voh, f0, f90);

    /* :232  (1.0 - fresnel) * distribution * geometric_shadowing * diffuse_colour */
    return v3_mul(v3_scale(v3_scale(v3_one_minus(fresnel), distribution), shadowing), m.diffuse_colour);
}

/* glam-pbr/src/lib.rs:248-256 */
v3 orc_refract(v3 incident, v3 normal, float ior) {
    float eta = 1.0f / ior;
    float n_dot_i = v3_dot(normal, incident);
    float k = 1.0f - eta * eta * (1.0f - n_dot_i * n_dot_i);
    return v3_sub(v3_scale(incident, eta), v3_scale(normal, eta * n_dot_i + sqrtf(k)));
}

/* glam-pbr/src/lib.rs:275-290 */
static v3 apply_volume_attenuation(v3 transmitted, float transmission_distance, float attenuation_distance,
                                   v3 attenuation_colour) {
    if (attenuation_distance == INFINITY) return transmitted;
    /* :285 -ln(colour) / distance ; :287 exp(-coefficient * distance) */
    v3 lnc = v3_new(logf(attenuation_colour.x), logf(attenuation_colour.y), logf(attenuation_colour.z)); /* :271-273 */
    v3 coeff = v3_divs(v3_neg(lnc), attenuation_distance);
    v3 e = v3_scale(v3_neg(coeff), transmission_distance);
    v3 transmittance = v3_new(expf(e.x), expf(e.y), expf(e.z));
    return v3_mul(transmittance, transmitted);
}

/* glam-pbr/src/lib.rs:292-354 */
v3 orc_ibl_volume_refraction(const orc_ibl_params* p, const orc_pyramid* fb, const orc_lut* lut) {
    orc_material_params m = p->material_params;
    /* :258-268 get_volume_transmission_ray */
    v3 refraction = orc_refract(v3_neg(p->view), p->normal, m.index_of_refraction);
    float ray_length = p->thickness * p->model_scale;
    v3 ray = v3_scale(v3_normalize(refraction), ray_length);
    v3 exit = v3_add(p->position, ray);

    v4 device = m4_mul_v4(&p->proj_view_matrix, v4_new(exit.x, exit.y, exit.z, 1.0f));
    float sx = device.x / device.w, sy = device.y / device.w;
    float tu = (sx + 1.0f) / 2.0f, tv = (sy + 1.0f) / 2.0f;

    /* :334-335; PerceptualRoughness::apply_ior :158-160 */
    float lod = log2f((float)p->framebuffer_size_x) *
                (m.perceptual_roughness * clampf(m.index_of_refraction * 2.0f - 2.0f, 0.0f, 1.0f));

    v3 transmitted = orc_sample_pyramid(fb, tu, tv, lod);
    v3 attenuated = apply_volume_attenuation(transmitted, ray_length, p->attenuation_distance, p->attenuation_colour);

    float nov = v3_dot(p->normal, p->view); /* :345 unclamped */
    v2 brdf = orc_sample_lut(lut, nov, m.perceptual_roughness);

    v3 f0 = orc_calculate_combined_f0(m);
    v3 f90 = orc_calculate_combined_f90(m);
    v3 specular_colour = v3_add(v3_scale(f0, brdf.x), v3_scale(f90, brdf.y));

    return v3_mul(v3_mul(v3_one_minus(specular_colour), attenuated), m.diffuse_colour);
}

/* ---- batch forms ---------------------------------------------------- */
static v3 from_tr3(tr_vec3 a) { return v3_new(a.x, a.y, a.z); }
static tr_vec3 to_tr3(v3 a) { tr_vec3 r = {a.x, a.y, a.z}; return r; }
static orc_material_params from_tr_mat(const tr_material_params* m) {
    orc_material_params r;
    r.diffuse_colour = from_tr3(m->diffuse_colour);
    r.metallic = m->metallic;
    r.perceptual_roughness = m->perceptual_roughness;
    r.index_of_refraction = m->index_of_refraction;
    r.specular_colour = from_tr3(m->specular_colour);
    r.specular_factor = m->specular_factor;
    return r;
}

void orc_eval_basic_brdf(uint32_t n, const tr_basic_brdf_params* p, tr_brdf_result* out) {
    for (uint32_t i = 0; i < n; i++) {
        orc_brdf_result r = orc_basic_brdf(from_tr3(p[i].normal), from_tr3(p[i].light), from_tr3(p[i].light_intensity),
                                           from_tr3(p[i].view), from_tr_mat(&p[i].material_params));
        out[i].diffuse = to_tr3(r.diffuse);
        out[i].specular = to_tr3(r.specular);
    }
}

void orc_eval_transmission_btdf(uint32_t n, const tr_transmission_btdf_params* p, tr_vec3* out) {
    for (uint32_t i = 0; i < n; i++)
        out[i] = to_tr3(orc_transmission_btdf(from_tr_mat(&p[i].material_params), from_tr3(p[i].normal),
                                              from_tr3(p[i].view), from_tr3(p[i].light)));
}

/* one iteration of evaluate_lights (lighting.rs:179-216) + evaluate_lights_transmission (:58-92) for a point light */
void orc_eval_point_light(uint32_t n, const tr_point_light_params* p, tr_point_light_result* out) {
    for (uint32_t i = 0; i < n; i++) {
        v3 direction;
        float distance, attenuation;
        orc_light_direction_and_attenuation(from_tr3(p[i].position), from_tr3(p[i].light_position), &direction, &distance,
                                            &attenuation);
        orc_material_params m = from_tr_mat(&p[i].material_params);
        v3 light = v3_scale(from_tr3(p[i].light_colour), attenuation);   /* light_emission = colour_emission * factor * attenuation */
        orc_brdf_result r = orc_basic_brdf(from_tr3(p[i].normal), direction, light, from_tr3(p[i].view), m);
        out[i].diffuse = to_tr3(r.diffuse);
        out[i].specular = to_tr3(r.specular);
        out[i].transmission = to_tr3(v3_mul(orc_transmission_btdf(m, from_tr3(p[i].normal), from_tr3(p[i].view), direction), light));
    }
}

void orc_eval_ibl_volume_refraction(uint32_t n, const tr_mat4* proj_view, const tr_ibl_volume_refraction_params* p,
                                    const orc_pyramid* fb, const orc_lut* lut, tr_vec3* out) {
    for (uint32_t i = 0; i < n; i++) {
        orc_ibl_params q;
        q.material_params = from_tr_mat(&p[i].material_params);
        q.framebuffer_size_x = p[i].framebuffer_size_x;
        q.normal = from_tr3(p[i].normal);
        q.view = from_tr3(p[i].view);
        memcpy(&q.proj_view_matrix, proj_view, sizeof(m4));
        q.position = from_tr3(p[i].position);
        q.thickness = p[i].thickness;
        q.model_scale = p[i].model_scale;
        q.attenuation_distance = p[i].attenuation_distance;
        q.attenuation_colour = from_tr3(p[i].attenuation_colour);
        out[i] = to_tr3(orc_ibl_volume_refraction(&q, fb, lut));
    }
}

/* oracle/oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp32, -ffp-contract=off) of the reference's
 * per-pixel light-transport path.  It exists to CHECK the CUDA kernels; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (transmission_renderer_b200/csrc)
 * never links, imports or falls back to anything in this directory.
 *
 * PARITY: PINNED TO THE REFERENCE'S OWN COMPILED CODE for every stage the reference
 * runs in a shader.  /root/reference has no tests or golden vectors and cannot be
 * built here (Rust -> SPIR-V, no cargo/rustc, no Vulkan ICD), but it ships its shaders
 * compiled: compiled-shaders/normal/<stage>.spv.  oracle/spv2c.py translates those modules to
 * C one SPIR-V instruction per statement, oracle/spv_harness.c runs them with the
 * reference's descriptor interface (oracle/build_ref.py -> oracle/_ref/libspvref.so), and
 * tests/test_reference_spirv.py requires this restatement to match them: bit-exact for
 * frustum_culling, demultiplex_draws, write_cluster_data, assign_lights_to_clusters (as
 * ascending sets), vertex_instanced_with_scale, fragment and fragment_transmission (fp32
 * pixels of whole frames), depth_pre_pass_alpha_clip (kill decisions), fragment_tonemap
 * (sRGB8).  The modules' outputs on the cases of tests/spirv_cases.py are committed as
 * tests/golden/spirv_golden.npz, so the check also runs where /root/reference is absent.
 * Also kept: known answers derived from the source text (tests/test_oracle_kat.py) and
 * an independent float64 numpy restatement (tests/ref_f64.py).
 * NOT pinned, because the reference used fixed-function hardware or host code there and
 * the definition is ours: the rasterisation rules
 * (raster.c), image sampling and RGBA16F stores, the blit filter of the mip chain
 * (mips.c), the ray/triangle arithmetic of ray queries (shadow.c), and colstodian's baked
 * tonemapper constants (caller input).  Those files say so.
 *
 * Third-party arithmetic absent from /root/reference and restated here from
 * its published behaviour: glam 0.19.0 (vector ops, vecmath.h), libm 0.2.1 /
 * GLSL.std.450 (sqrt/pow/ln/exp/log2/cos/sin -> glibc libm), the Vulkan
 * implementation's texture filtering, blit and RGBA16F store
 * (SURVEY.md Appendix E -> sampler/mip code in shade.c / mips.c).
 */
#ifndef ORACLE_H
#define ORACLE_H

#include "../include/tr_abi.h"
#include "vecmath.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- glam-pbr (glam-pbr/src/lib.rs) -------------------------------- */
typedef struct {
    v3 diffuse_colour;
    float metallic;
    float perceptual_roughness;
    float index_of_refraction;
    v3 specular_colour;
    float specular_factor;
} orc_material_params; /* glam-pbr/src/lib.rs:171-179 */

typedef struct { v3 diffuse, specular; } orc_brdf_result; /* :437-441 */

void orc_light_direction_and_attenuation(v3 fragment_position, v3 light_position, v3* direction,
                                         float* distance, float* attenuation);                 /* :12-23 */
float orc_d_ggx(float noh, float alpha);                                                       /* :101-109 */
float orc_v_smith_ggx_correlated(float nov, float nol, float alpha);                           /* :114-133 */
v3 orc_fresnel_schlick(float vdoth, v3 f0, v3 f90);                                            /* :137-139 */
float orc_ior_to_dielectric_f0(float ior);                                                     /* :192-195 */
v3 orc_calculate_combined_f0(orc_material_params m);                                           /* :425-430 */
v3 orc_calculate_combined_f90(orc_material_params m);                                          /* :432-435 */
orc_brdf_result orc_basic_brdf(v3 normal, v3 light, v3 light_intensity, v3 view, orc_material_params m); /* :377-423 */
v3 orc_transmission_btdf(orc_material_params m, v3 normal, v3 view, v3 light);                 /* :200-233 */
v3 orc_refract(v3 incident, v3 normal, float ior);                                             /* :248-256 */

/* Sampled images */
typedef struct {
    uint32_t levels;
    uint32_t width[16], height[16];
    const uint16_t* data[16]; /* RGBA16F bits per level */
} orc_pyramid;
typedef struct {
    uint32_t width, height;
    const uint8_t* rgba8;
} orc_lut;

v3 orc_sample_pyramid(const orc_pyramid* p, float u, float v, float lod);  /* sample_by_lod, shader lib.rs:135-138 */
v2 orc_sample_lut(const orc_lut* lut, float n_dot_v, float roughness);     /* shader lib.rs:126-133 */

typedef struct {
    orc_material_params material_params;
    uint32_t framebuffer_size_x;
    v3 normal, view;
    m4 proj_view_matrix;
    v3 position;
    float thickness, model_scale, attenuation_distance;
    v3 attenuation_colour;
} orc_ibl_params; /* glam-pbr/src/lib.rs:235-246 */
v3 orc_ibl_volume_refraction(const orc_ibl_params* p, const orc_pyramid* fb, const orc_lut* lut); /* :292-354 */

/* batch forms mirroring tr_eval_* in include/tr_abi.h */
void orc_eval_basic_brdf(uint32_t n, const tr_basic_brdf_params* p, tr_brdf_result* out);
void orc_eval_transmission_btdf(uint32_t n, const tr_transmission_btdf_params* p, tr_vec3* out);
void orc_eval_point_light(uint32_t n, const tr_point_light_params* p, tr_point_light_result* out);
void orc_eval_ibl_volume_refraction(uint32_t n, const tr_mat4* proj_view, const tr_ibl_volume_refraction_params* p,
                                    const orc_pyramid* fb, const orc_lut* lut, tr_vec3* out);

/* ---- shared-structs methods ---------------------------------------- */
float orc_log2_spec(float x);                                             /* our fp32 log2 definition */
float orc_linear_depth(const tr_light_cluster_coefficients* c, float d);  /* shared-structs:54-58 */
uint32_t orc_get_depth_slice(const tr_light_cluster_coefficients* c, float d); /* :61-63 */
float orc_slice_to_depth(const tr_light_cluster_coefficients* c, uint32_t slice); /* :65-67 */
void orc_light_cluster_coefficients_new(float z_near, float z_far, uint32_t slices, tr_light_cluster_coefficients* out); /* :44-52 */
float orc_spotlight_factor(const tr_light* l, v3 direction_to_light);     /* :129-138 */
float orc_aabb_distance_sq(const tr_cluster_aabb* a, v3 p);               /* :291-298 */
int orc_aabb_cull_spotlight(const tr_cluster_aabb* a, v3 origin, v3 direction, float angle, float range); /* :301-319 */
void orc_mat4_inverse(const tr_mat4* m, tr_mat4* out);                    /* double-precision cofactor inverse, rounded to f32 */

/* ---- compute entry points (shader/src/lib.rs) ----------------------- */
int orc_cull(tr_vec4 sphere, const tr_packed_similarity* t, const tr_culling_push_constants* pc); /* :442-469, 1 = culled */
/* frustum_culling + demultiplex_draws with deterministic (ascending) order, :411-517 */
void orc_frustum_culling(const tr_instance* inst, uint32_t n_inst, const tr_primitive_info* prims, uint32_t n_prims,
                         const tr_culling_push_constants* pc, uint32_t* instance_counts /*[n_prims]*/,
                         uint32_t* visible_ids /*[n_inst]*/, uint32_t* n_visible);
void orc_demultiplex_draws(const tr_primitive_info* prims, const uint32_t* instance_counts, uint32_t n_prims,
                           tr_draw_indexed_indirect_command* draws[4] /* each [n_prims] */, uint32_t draw_counts[4]);
void orc_write_cluster_data(const tr_uniforms* u, const tr_write_cluster_data_push_constants* pc, uint32_t nz,
                            tr_cluster_aabb* out /*[nx*ny*nz]*/);           /* :519-594 */
void orc_assign_lights_to_clusters(const tr_light* lights, uint32_t n_lights, const tr_cluster_aabb* clusters,
                                   uint32_t n_clusters, const tr_assign_lights_push_constants* pc,
                                   uint32_t* counts /*[n_clusters]*/, uint32_t* indices /*[n_clusters*128]*/); /* :596-645 */

/* ---- fragment stage -------------------------------------------------- */
typedef struct {
    uint32_t width, height;
    const float* depth;
    const float* normal;
    const float* uv;
    const uint32_t* material_id;
    const float* scale;    /* may be NULL */
    const float* position; /* may be NULL => reconstruct from depth */
    const float* duv;      /* [h*w*4] uv(x+1,y)-uv(x,y), uv(x,y+1)-uv(x,y) on the pixel's own triangle; may be NULL (=> 0) */
    const float* ddepth;   /* [h*w*2] the same forward differences of frag_coord.z; may be NULL (=> 0) */
} orc_gbuffer;

/* Bindless sampled images of the material system (descriptor set 0 binding 0, src/descriptor_sets.rs; loader
 * src/model_loading.rs:340-379): RGBA8, sRGB or UNORM, with the full mip chain the loader creates. */
typedef struct {
    uint32_t width, height, levels, srgb;
    const uint8_t* data[16]; /* RGBA8 per level, level k is max(1, w >> k) x max(1, h >> k) */
} orc_texture;
/* texture.sample(sampler, uv) with the repeat sampler (linear min/mag/mip, src/main.rs:683-692) and the level of detail
 * from the uv differences to the right and lower neighbour (Vulkan: lambda = log2(max(|d(uv*size)/dx|, |d(uv*size)/dy|))) */
v4 orc_sample_texture(const orc_texture* t, v2 uv, v2 duv_dx, v2 duv_dy);
float orc_srgb8_to_linear(uint8_t c);

/* ---- ray-queried shadows (oracle/shadow.c; lighting.rs:97-125, src/acceleration_structures.rs) ---- */
typedef struct orc_mesh_s orc_mesh;
typedef struct orc_accel orc_accel;
orc_accel* orc_accel_build(const orc_mesh* mesh, const tr_instance* inst, uint32_t n_inst, const tr_primitive_info* prims,
                           uint32_t n_prims);
void orc_accel_free(orc_accel* a);
uint32_t orc_accel_instance_count(const orc_accel* a); /* instances in the top-level set (draw_buffer_index < 2) */
float orc_trace_shadow(const orc_accel* a, v3 origin, v3 direction, float t_max);       /* 1 lit, 0 occluded */
float orc_trace_shadow_brute(const orc_accel* a, v3 origin, v3 direction, float t_max); /* same, no trees */
void orc_trace_shadow_rays(const orc_accel* a, uint32_t n, const float* origins, const float* directions, const float* t_max,
                           int brute, uint8_t* lit);
int orc_slab_test(const float origin[3], const float direction[3], float t_min, float t_max, const float lo[3], const float hi[3]);

typedef struct {
    const tr_push_constants* pc;
    const tr_uniforms* uniforms;
    const tr_material_info* materials;
    uint32_t n_materials;
    const tr_light* lights;
    uint32_t n_lights;
    const uint32_t* cluster_light_counts;
    const uint32_t* cluster_light_indices;
    uint32_t n_clusters;
    const orc_texture* textures; /* may be NULL when no material binds one */
    uint32_t n_textures;
    const orc_accel* accel; /* NULL: no ray queries (acceleration_structure_address == 0) */
} orc_scene;

/* what the 2x2 quad gave the reference's fragment stage implicitly: differences to the right / lower neighbour */
typedef struct {
    v2 duv_dx, duv_dy;
    v3 dpos_dx, dpos_dy; /* world position differences (ddx / ddy of -view_vector, lighting.rs:246-249) */
} orc_frag_derivatives;

/* `fragment`, shader/src/lib.rs:164-249, for one pixel; returns rgba */
v4 orc_fragment(v3 position, v3 normal, v2 uv, uint32_t material_id, v4 frag_coord, const orc_scene* s,
                const orc_frag_derivatives* d /* NULL => zero */);
/* `fragment_transmission`, shader/src/lib.rs:37-162 */
v4 orc_fragment_transmission(v3 position, v3 normal, v2 uv, uint32_t material_id, float model_scale, v4 frag_coord,
                             const orc_scene* s, const orc_pyramid* fb, const orc_lut* lut, const orc_frag_derivatives* d);

/* Frame drivers: rows [y0,y1), OpenMP over rows.  Outputs are full-frame
 * buffers; any may be NULL.  Opaque: clears to (0,0,0,1) then shades covered
 * pixels, writing the same value to hdr and to the sampled target
 * (lib.rs:247-248).  Transmission: LOADs hdr and overwrites covered pixels. */
void orc_shade_opaque_frame(const orc_gbuffer* g, const orc_scene* s, uint32_t y0, uint32_t y1, float* hdr_f32,
                            uint16_t* hdr_f16, uint16_t* opaque_f16);
void orc_shade_transmission_frame(const orc_gbuffer* g, const orc_scene* s, const orc_pyramid* fb, const orc_lut* lut,
                                  uint32_t y0, uint32_t y1, float* hdr_f32, uint16_t* hdr_f16);
/* The same frame loop with the per-fragment function supplied by the caller (oracle/spv_harness.c runs the reference's
 * shipped SPIR-V through it): identical G-buffer decode, derivative decode and stores. transmissive = 0: `fragment`
 * semantics (clear + two targets), 1: `fragment_transmission` semantics (LOAD, covered pixels only). */
typedef v4 (*orc_fragment_fn)(v3 position, v3 normal, v2 uv, uint32_t material_id, float model_scale, v4 frag_coord,
                              const orc_frag_derivatives* d, void* user);
void orc_shade_frame_with(const orc_gbuffer* g, const tr_push_constants* pc, int transmissive, uint32_t y0, uint32_t y1,
                          orc_fragment_fn fn, void* user, float* hdr_f32, uint16_t* hdr_f16, uint16_t* opaque_f16);
/* occluded-ray bits of a layer: [5][h*w], planes 0-3 = cluster-list position, plane 4 bit 0 = sun */
void orc_shadow_mask_frame(const orc_gbuffer* g, const orc_scene* s, uint32_t y0, uint32_t y1, uint32_t* mask);

/* ---- mip chain (src/main.rs:2054-2063, 2590-2592; Appendix E) -------- */
uint32_t orc_mip_levels_for_size(uint32_t w, uint32_t h);
void orc_mip_size(uint32_t w, uint32_t h, uint32_t level, uint32_t* lw, uint32_t* lh);
void orc_downsample_level(const uint16_t* src, uint32_t sw, uint32_t sh, uint16_t* dst, uint32_t dw, uint32_t dh);
void orc_f32_to_f16(const float* in, uint16_t* out, size_t n);
void orc_f16_to_f32(const uint16_t* in, float* out, size_t n);

/* ---- tonemap (shader/src/tonemapping.rs, lib.rs:683-697) ------------- */
v3 orc_lottes_tonemap(v3 color, const tr_baked_lottes_tonemapper_params* p);
uint8_t orc_srgb8_encode(float linear01);
void orc_tonemap_frame(const uint16_t* hdr_f16, uint32_t w, uint32_t h, uint32_t y0, uint32_t y1,
                       const tr_baked_lottes_tonemapper_params* p, uint8_t* rgba8);

/* ---- visibility (software stand-in for the rasteriser) ---- */
struct orc_mesh_s {
    const float* positions; /* [n_vertices*3] */
    const float* normals;   /* [n_vertices*3] */
    const float* uvs;       /* [n_vertices*2] */
    const uint32_t* indices;
    uint32_t n_vertices, n_indices;
};
void orc_visibility(const orc_mesh* mesh, const tr_instance* inst, uint32_t n_inst, const tr_primitive_info* prims,
                    uint32_t n_prims, const uint32_t* visible_ids, uint32_t n_visible, const tr_push_constants* pc,
                    uint32_t y0, uint32_t y1, float* depth0, float* normal0, float* uv0, uint32_t* mat0, float* depth1,
                    float* normal1, float* uv1, uint32_t* mat1, float* scale1,
                    float* duv0, float* ddepth0, float* duv1, float* ddepth1 /* derivative planes, any may be NULL */,
                    const tr_material_info* materials /* NULL: no alpha clipping */, const orc_texture* textures,
                    uint32_t n_textures);

void orc_vertex_instanced_with_scale(v3 position, v3 normal, const tr_instance* instance, const tr_push_constants* pc,
                                     v4* clip, v3* out_position, v3* out_normal, uint32_t* out_material_id, float* out_scale);
int orc_alpha_clip_kills(const tr_material_info* m, const orc_texture* textures, uint32_t n_textures, v2 uv, v2 duv_dx,
                         v2 duv_dy);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif

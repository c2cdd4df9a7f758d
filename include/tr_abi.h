/* tr_abi.h — the drop-in boundary of the B200-native light-transport path.
 *
 * Plain C ABI (extern "C", pointers + sizes, no torch/CUDA types in the
 * signatures).  Two things live here:
 *
 *  1. the reference's `#[repr(C)]` host<->GPU structs, byte for byte
 *     (/root/reference/shared-structs/src/lib.rs, offsets verified against the
 *     shipped SPIR-V) — every size/offset is pinned by a
 *     static assert below;
 *  2. the entry points that replace the reference's Vulkan plumbing
 *     (src/render_passes.rs, src/pipelines.rs, src/descriptor_sets.rs) and the
 *     frame recorder `record()` (src/main.rs:1551-2263).  Each export cites
 *     the reference interface it stands in for.
 *
 * All functions return 0 (TR_OK) or a negative tr_status; tr_last_error()
 * gives a thread-local message.  Nothing here ever falls back to a CPU path:
 * if the CUDA device or a kernel is unavailable the call fails.
 */
#ifndef TR_ABI_H
#define TR_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#define TR_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define TR_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

#define TR_ALIGN16 __attribute__((aligned(16)))
#if defined(_WIN32)
#define TR_API
#else
#define TR_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------ */
/* glam 0.19 value types (SSE2 layout: Vec4/Quat/Vec3A/Mat4 16-aligned) */
/* ------------------------------------------------------------------ */
typedef struct { float x, y; } tr_vec2;
typedef struct { uint32_t x, y; } tr_uvec2;
typedef struct { float x, y, z; } tr_vec3;
typedef struct TR_ALIGN16 { float x, y, z, w; } tr_vec4;
typedef struct TR_ALIGN16 { float x, y, z, _pad; } tr_vec3a;
typedef struct TR_ALIGN16 { float x, y, z, w; } tr_quat;
typedef struct TR_ALIGN16 { tr_vec4 col[4]; } tr_mat4; /* column-major */

/* ------------------------------------------------------------------ */
/* shared-structs (reference: shared-structs/src/lib.rs)               */
/* ------------------------------------------------------------------ */

/* shared-structs/src/lib.rs:11-16 */
typedef struct TR_ALIGN16 {
    tr_mat4 proj_view;
    tr_vec3a view_position;
    tr_uvec2 framebuffer_size;
    uint64_t acceleration_structure_address; /* 0 = no ray queries, else the handle of tr_build_acceleration_structures */
} tr_push_constants;

/* shared-structs/src/lib.rs:35-41 */
typedef struct {
    float z_near, z_far, scale, bias;
    uint32_t num_depth_slices;
} tr_light_cluster_coefficients;

/* shared-structs/src/lib.rs:21-29 */
typedef struct TR_ALIGN16 {
    tr_light_cluster_coefficients light_clustering_coefficients;
    tr_vec3a sun_dir;
    tr_vec3a sun_intensity;
    tr_vec2 cluster_size_in_pixels;
    tr_uvec2 num_clusters;
    uint32_t debug_clusters;         /* must be 0 */
    uint32_t ggx_lut_texture_index;  /* kept for layout; the LUT is bound by tr_set_ggx_lut */
} tr_uniforms;

/* shared-structs/src/lib.rs:74-78 */
typedef struct TR_ALIGN16 {
    tr_vec4 position_and_spotlight_epsilon;
    tr_vec4 colour_emission_and_falloff_distance_sq;
    tr_vec4 spotlight_direction_and_outer_angle; /* w == 0 => point light */
} tr_light;

/* shared-structs/src/lib.rs:143-153 */
typedef struct {
    int32_t diffuse, metallic_roughness, normal_map, emissive, occlusion,
        transmission, thickness, specular, specular_colour; /* -1 = none */
} tr_textures;

/* shared-structs/src/lib.rs:157-173 */
typedef struct TR_ALIGN16 {
    tr_textures textures;
    float metallic_factor;
    float roughness_factor;
    float alpha_clipping_cutoff;
    tr_vec4 diffuse_factor;
    tr_vec3a emissive_factor;
    float normal_map_scale;
    float occlusion_strength;
    float index_of_refraction;
    float transmission_factor;
    float thickness_factor;
    float attenuation_distance;
    tr_vec3a attenuation_colour;
    float specular_factor;
    tr_vec3a specular_colour_factor;
} tr_material_info;

/* shared-structs/src/lib.rs:178-181 */
typedef struct TR_ALIGN16 {
    tr_vec4 translation_and_scale; /* w = uniform scale */
    tr_quat rotation;
} tr_packed_similarity;

/* shared-structs/src/lib.rs:253-257 */
typedef struct TR_ALIGN16 {
    tr_packed_similarity transform;
    uint32_t primitive_id;
    uint32_t material_id;
} tr_instance;

/* shared-structs/src/lib.rs:262-268 */
typedef struct TR_ALIGN16 {
    tr_vec4 packed_bounding_sphere; /* xyz centre, w radius, model space */
    uint32_t draw_buffer_index;     /* 0 opaque, 1 alpha clip, 2 transmission, 3 transmission alpha clip (model_loading.rs:68-78) */
    uint32_t index_count;
    uint32_t first_index;
    uint32_t first_instance;
} tr_primitive_info;

/* shared-structs/src/lib.rs:273-280 */
typedef struct TR_ALIGN16 {
    tr_mat4 view;
    tr_vec2 frustum_x_xz;
    tr_vec2 frustum_y_yz;
    float z_near;
} tr_culling_push_constants;

/* shared-structs/src/lib.rs:285-288 */
typedef struct TR_ALIGN16 {
    tr_vec3a min;
    tr_vec3a max;
} tr_cluster_aabb;

/* shared-structs/src/lib.rs:336-339 */
typedef struct TR_ALIGN16 {
    tr_mat4 inverse_perspective;
    tr_uvec2 screen_dimensions;
} tr_write_cluster_data_push_constants;

/* shared-structs/src/lib.rs:344-347 */
typedef struct TR_ALIGN16 {
    tr_mat4 view_matrix;
    tr_quat view_rotation;
} tr_assign_lights_push_constants;

/* shader/src/lib.rs:401-409 (vk::DrawIndexedIndirectCommand) */
typedef struct {
    uint32_t index_count;
    uint32_t instance_count;
    uint32_t first_index;
    int32_t vertex_offset;
    uint32_t first_instance;
} tr_draw_indexed_indirect_command;

/* shader/src/tonemapping.rs:28-38 */
typedef struct {
    float a, b, c, d, crosstalk, saturation, cross_saturation;
} tr_baked_lottes_tonemapper_params;

#define TR_MAX_LIGHTS_PER_CLUSTER 128u /* shared-structs/src/lib.rs:322 */
#define TR_NUM_DRAW_BUFFERS 4u         /* shader/src/lib.rs:471 */

TR_STATIC_ASSERT(sizeof(tr_push_constants) == 96, "PushConstants");
TR_STATIC_ASSERT(offsetof(tr_push_constants, view_position) == 64, "PushConstants.view_position");
TR_STATIC_ASSERT(offsetof(tr_push_constants, framebuffer_size) == 80, "PushConstants.framebuffer_size");
TR_STATIC_ASSERT(offsetof(tr_push_constants, acceleration_structure_address) == 88, "PushConstants.as_address");
TR_STATIC_ASSERT(sizeof(tr_light_cluster_coefficients) == 20, "LightClusterCoefficients");
TR_STATIC_ASSERT(sizeof(tr_uniforms) == 96, "Uniforms");
TR_STATIC_ASSERT(offsetof(tr_uniforms, sun_dir) == 32, "Uniforms.sun_dir");
TR_STATIC_ASSERT(offsetof(tr_uniforms, sun_intensity) == 48, "Uniforms.sun_intensity");
TR_STATIC_ASSERT(offsetof(tr_uniforms, cluster_size_in_pixels) == 64, "Uniforms.cluster_size_in_pixels");
TR_STATIC_ASSERT(offsetof(tr_uniforms, num_clusters) == 72, "Uniforms.num_clusters");
TR_STATIC_ASSERT(offsetof(tr_uniforms, debug_clusters) == 80, "Uniforms.debug_clusters");
TR_STATIC_ASSERT(offsetof(tr_uniforms, ggx_lut_texture_index) == 84, "Uniforms.ggx_lut_texture_index");
TR_STATIC_ASSERT(sizeof(tr_light) == 48, "Light");
TR_STATIC_ASSERT(sizeof(tr_textures) == 36, "Textures");
TR_STATIC_ASSERT(sizeof(tr_material_info) == 160, "MaterialInfo");
TR_STATIC_ASSERT(offsetof(tr_material_info, metallic_factor) == 36, "MaterialInfo.metallic_factor");
TR_STATIC_ASSERT(offsetof(tr_material_info, roughness_factor) == 40, "MaterialInfo.roughness_factor");
TR_STATIC_ASSERT(offsetof(tr_material_info, alpha_clipping_cutoff) == 44, "MaterialInfo.alpha_clipping_cutoff");
TR_STATIC_ASSERT(offsetof(tr_material_info, diffuse_factor) == 48, "MaterialInfo.diffuse_factor");
TR_STATIC_ASSERT(offsetof(tr_material_info, emissive_factor) == 64, "MaterialInfo.emissive_factor");
TR_STATIC_ASSERT(offsetof(tr_material_info, normal_map_scale) == 80, "MaterialInfo.normal_map_scale");
TR_STATIC_ASSERT(offsetof(tr_material_info, occlusion_strength) == 84, "MaterialInfo.occlusion_strength");
TR_STATIC_ASSERT(offsetof(tr_material_info, index_of_refraction) == 88, "MaterialInfo.index_of_refraction");
TR_STATIC_ASSERT(offsetof(tr_material_info, transmission_factor) == 92, "MaterialInfo.transmission_factor");
TR_STATIC_ASSERT(offsetof(tr_material_info, thickness_factor) == 96, "MaterialInfo.thickness_factor");
TR_STATIC_ASSERT(offsetof(tr_material_info, attenuation_distance) == 100, "MaterialInfo.attenuation_distance");
TR_STATIC_ASSERT(offsetof(tr_material_info, attenuation_colour) == 112, "MaterialInfo.attenuation_colour");
TR_STATIC_ASSERT(offsetof(tr_material_info, specular_factor) == 128, "MaterialInfo.specular_factor");
TR_STATIC_ASSERT(offsetof(tr_material_info, specular_colour_factor) == 144, "MaterialInfo.specular_colour_factor");
TR_STATIC_ASSERT(sizeof(tr_packed_similarity) == 32, "PackedSimilarity");
TR_STATIC_ASSERT(sizeof(tr_instance) == 48, "Instance");
TR_STATIC_ASSERT(offsetof(tr_instance, primitive_id) == 32, "Instance.primitive_id");
TR_STATIC_ASSERT(offsetof(tr_instance, material_id) == 36, "Instance.material_id");
TR_STATIC_ASSERT(sizeof(tr_primitive_info) == 32, "PrimitiveInfo");
TR_STATIC_ASSERT(offsetof(tr_primitive_info, draw_buffer_index) == 16, "PrimitiveInfo.draw_buffer_index");
TR_STATIC_ASSERT(sizeof(tr_culling_push_constants) == 96, "CullingPushConstants");
TR_STATIC_ASSERT(offsetof(tr_culling_push_constants, frustum_x_xz) == 64, "CullingPushConstants.frustum_x_xz");
TR_STATIC_ASSERT(offsetof(tr_culling_push_constants, frustum_y_yz) == 72, "CullingPushConstants.frustum_y_yz");
TR_STATIC_ASSERT(offsetof(tr_culling_push_constants, z_near) == 80, "CullingPushConstants.z_near");
TR_STATIC_ASSERT(sizeof(tr_cluster_aabb) == 32, "ClusterAabb");
TR_STATIC_ASSERT(sizeof(tr_write_cluster_data_push_constants) == 80, "WriteClusterDataPushConstants");
TR_STATIC_ASSERT(offsetof(tr_write_cluster_data_push_constants, screen_dimensions) == 64, "WriteClusterData.screen_dimensions");
TR_STATIC_ASSERT(sizeof(tr_assign_lights_push_constants) == 80, "AssignLightsPushConstants");
TR_STATIC_ASSERT(offsetof(tr_assign_lights_push_constants, view_rotation) == 64, "AssignLights.view_rotation");
TR_STATIC_ASSERT(sizeof(tr_draw_indexed_indirect_command) == 20, "DrawIndexedIndirectCommand");
TR_STATIC_ASSERT(sizeof(tr_baked_lottes_tonemapper_params) == 28, "BakedLottesTonemapperParams");

/* ------------------------------------------------------------------ */
/* glam-pbr batch contracts (reference: glam-pbr/src/lib.rs)           */
/* SoA-free "array of params" form of the pub fns; Rust closures cannot */
/* cross a C ABI so the samplers are data (pyramid + LUT bound on ctx). */
/* ------------------------------------------------------------------ */

/* glam-pbr/src/lib.rs:171-179 MaterialParams */
typedef struct {
    tr_vec3 diffuse_colour;
    float metallic;
    float perceptual_roughness;
    float index_of_refraction;
    tr_vec3 specular_colour;
    float specular_factor;
} tr_material_params;

/* glam-pbr/src/lib.rs:163-169 BasicBrdfParams */
typedef struct {
    tr_vec3 normal;
    tr_vec3 light;
    tr_vec3 light_intensity;
    tr_vec3 view;
    tr_material_params material_params;
} tr_basic_brdf_params;

/* glam-pbr/src/lib.rs:437-441 BrdfResult */
typedef struct {
    tr_vec3 diffuse;
    tr_vec3 specular;
} tr_brdf_result;

/* glam-pbr/src/lib.rs:200-205 arguments of transmission_btdf */
typedef struct {
    tr_material_params material_params;
    tr_vec3 normal;
    tr_vec3 view;
    tr_vec3 light;
} tr_transmission_btdf_params;

/* glam-pbr/src/lib.rs:235-246 IblVolumeRefractionParams (proj_view_matrix passed once per batch) */
typedef struct {
    tr_material_params material_params;
    uint32_t framebuffer_size_x;
    tr_vec3 normal;
    tr_vec3 view;
    tr_vec3 position;
    float thickness;
    float model_scale;
    float attenuation_distance;
    tr_vec3 attenuation_colour;
} tr_ibl_volume_refraction_params;

/* One iteration of the clustered-light loops for a point light — shader/src/lighting.rs:179-216 (`evaluate_lights`) and
 * :58-92 (`evaluate_lights_transmission`): light_direction_and_attenuation(position, light_position) (glam-pbr
 * lib.rs:12-23), then basic_brdf(normal, direction, colour * attenuation, view, material) and
 * transmission_btdf(material, normal, view, direction) * colour * attenuation.  `normal` and `view` are unit vectors
 * (the shaders normalise them before the loop).  tr_eval_point_light runs the SAME device code as the frame kernels'
 * light loop (fast regime + adaptive exactness), so the hot loop has a per-element contract test of its own. */
typedef struct {
    tr_vec3 normal;
    tr_vec3 view;
    tr_vec3 position;       /* fragment, world space */
    tr_vec3 light_position;
    tr_vec3 light_colour;   /* Light::colour_emission (rgb = colour * intensity) */
    tr_material_params material_params;
} tr_point_light_params;

typedef struct {
    tr_vec3 diffuse;        /* BrdfResult.diffuse */
    tr_vec3 specular;       /* BrdfResult.specular */
    tr_vec3 transmission;   /* transmission_btdf(..) * light */
} tr_point_light_result;

TR_STATIC_ASSERT(sizeof(tr_material_params) == 40, "MaterialParams (C form)");
TR_STATIC_ASSERT(sizeof(tr_point_light_params) == 100, "point-light loop iteration params");
TR_STATIC_ASSERT(sizeof(tr_point_light_result) == 36, "point-light loop iteration result");
TR_STATIC_ASSERT(sizeof(tr_basic_brdf_params) == 88, "BasicBrdfParams (C form)");
TR_STATIC_ASSERT(sizeof(tr_transmission_btdf_params) == 76, "transmission_btdf params (C form)");
TR_STATIC_ASSERT(sizeof(tr_ibl_volume_refraction_params) == 104, "IblVolumeRefractionParams (C form)");

/* ------------------------------------------------------------------ */
/* Status, config, opaque context                                       */
/* ------------------------------------------------------------------ */
typedef enum {
    TR_OK = 0,
    TR_ERR_INVALID_ARG = -1,
    TR_ERR_UNSUPPORTED = -2, /* e.g. debug_clusters != 0 */
    TR_ERR_CUDA = -3,
    TR_ERR_NCCL = -4,
    TR_ERR_OOM = -5,
    TR_ERR_STATE = -6 /* call order violated (e.g. shade before a G-buffer exists) */
} tr_status;

typedef struct tr_ctx tr_ctx;

typedef struct {
    uint32_t width, height;  /* full framebuffer size (main.rs: extent) */
    int32_t device;          /* CUDA ordinal */
    uint32_t band_y0, band_y1; /* rows this context shades; y1 == 0 => whole frame */
    uint32_t flags;          /* TR_FLAG_* */
} tr_config;

#define TR_FLAG_HDR_F32_DEBUG 1u /* also keep an fp32 copy of the HDR target (parity tests) */

/* G-buffer planes (structure-of-arrays, one element per pixel, row-major,
 * row stride = width).  This is what the reference's rasteriser hands the
 * fragment stage as varyings (shader/src/lib.rs:37-56,164-181), flattened:
 *   depth        frag_coord.z (reversed-Z, 0 = empty pixel; main.rs:1586-1591)
 *   normal       interpolated `rotation * normal` (un-normalised, lib.rs:356)
 *   uv           interpolated uv (lib.rs:357)
 *   material_id  flat (lib.rs:358)
 *   scale        flat similarity.scale (transmissive layer only, lib.rs:388)
 *   position     optional explicit world position; NULL => reconstructed from
 *                depth and the pixel centre with inverse(proj_view)
 */
typedef struct {
    const float* depth;          /* [h*w]   */
    const float* normal;         /* [h*w*3] */
    const float* uv;             /* [h*w*2] */
    const uint32_t* material_id; /* [h*w]   */
    const float* scale;          /* [h*w] or NULL (layer 0) */
    const float* position;       /* [h*w*3] or NULL */
    /* forward differences to the right / lower neighbour on the pixel's own triangle: what the 2x2 quad gives the
     * reference's fragment stage implicitly (texture level of detail; ddx/ddy in lighting.rs:246-249).  NULL => 0. */
    const float* duv;            /* [h*w*4] du/dx, dv/dx, du/dy, dv/dy */
    const float* ddepth;         /* [h*w*2] d(frag_coord.z)/dx, /dy */
} tr_gbuffer_planes;

typedef struct {
    float* depth;
    float* normal;
    float* uv;
    uint32_t* material_id;
    float* scale;
    float* position; /* filled only if the layer carries explicit positions */
    float* duv;      /* filled only once textures are bound (tr_set_texture) */
    float* ddepth;
} tr_gbuffer_planes_out;

enum { TR_LAYER_OPAQUE = 0, TR_LAYER_TRANSMISSIVE = 1 };

/* ------------------------------------------------------------------ */
/* Lifecycle — replaces Pipelines::new / DescriptorSets::allocate       */
/* (src/pipelines.rs:46-172, src/descriptor_sets.rs:150-316) and the    */
/* LoopDestroyed cleanup (src/main.rs:1416-1445).                        */
/* ------------------------------------------------------------------ */
TR_API int32_t tr_create(const tr_config* config, tr_ctx** out_ctx);
TR_API int32_t tr_destroy(tr_ctx* ctx);
/* swapchain-resize path, src/main.rs:996-1166: reallocates depth/hdr/pyramid. */
TR_API int32_t tr_resize(tr_ctx* ctx, uint32_t width, uint32_t height);
TR_API int32_t tr_set_band(tr_ctx* ctx, uint32_t y0, uint32_t y1);
TR_API const char* tr_last_error(void);
TR_API const char* tr_version(void);
/* Run every subsequent call on this CUDA stream (a cudaStream_t cast to void*; NULL = the context's own). */
TR_API int32_t tr_set_stream(tr_ctx* ctx, void* cuda_stream);
/* fence wait, src/main.rs:1290 */
TR_API int32_t tr_sync(tr_ctx* ctx);

/* ------------------------------------------------------------------ */
/* Uploads — replace the descriptor writes, src/main.rs:715-828         */
/* ------------------------------------------------------------------ */
TR_API int32_t tr_set_instances(tr_ctx* ctx, const tr_instance* instances, uint32_t n);      /* set1/b0, main.rs:2527 */
TR_API int32_t tr_set_primitives(tr_ctx* ctx, const tr_primitive_info* prims, uint32_t n);   /* set0/b7 + cull set/b0 */
TR_API int32_t tr_set_materials(tr_ctx* ctx, const tr_material_info* materials, uint32_t n); /* set0/b2 */
TR_API int32_t tr_set_lights(tr_ctx* ctx, const tr_light* lights, uint32_t n);               /* set2/b0, main.rs:472-478 */
TR_API int32_t tr_set_uniforms(tr_ctx* ctx, const tr_uniforms* uniforms);                    /* set0/b3, main.rs:1231 */
/* set0/b0[ggx_lut_texture_index]: RGBA8 UNORM, one mip (main.rs:295-330). */
TR_API int32_t tr_set_ggx_lut(tr_ctx* ctx, const uint8_t* rgba8, uint32_t width, uint32_t height);
/* set0/b0[index]: one bindless sampled image of the material system (src/descriptor_sets.rs, MAX_IMAGES = 193,
 * src/main.rs:59), as the loader creates it (src/model_loading.rs:340-379): RGBA8, `srgb` != 0 for R8G8B8A8_SRGB, with
 * its mip chain.  levels[k] points to mip k, which is max(1, width >> k) x max(1, height >> k) texels, tightly packed.
 * Sampled with the repeat sampler (linear min/mag/mip, src/main.rs:683-692). */
#define TR_MAX_IMAGES 193u
TR_API int32_t tr_set_texture(tr_ctx* ctx, uint32_t index, const uint8_t* const* levels, uint32_t n_levels,
                              uint32_t width, uint32_t height, int32_t srgb);
/* vertex bindings 0..2 + index buffer (pipelines.rs:291-307, main.rs:2516-2557). */
TR_API int32_t tr_set_mesh(tr_ctx* ctx, const float* positions, const float* normals, const float* uvs,
                           uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices);

/* ------------------------------------------------------------------ */
/* Ray-queried shadows — the `--ray-tracing` option (src/main.rs:84,   */
/* 577-658) and the RayQueryKHR variant of the fragment shaders          */
/* (shader/src/lighting.rs:22-32, 64-71, 97-125, 154-165, 186-195).      */
/* ------------------------------------------------------------------ */
/* build_acceleration_structures_from_primitives + build_top_level_acceleration_structure_from_instances
 * (src/acceleration_structures.rs:6-186, called at src/main.rs:594-649): one bottom-level structure per primitive
 * over the bound mesh, one top-level structure over the bound instances whose primitive has draw_buffer_index < 2
 * (src/main.rs:614-625).  *address is what the reference puts in PushConstants.acceleration_structure_address
 * (src/main.rs:856-859); passing it to tr_frame / tr_shade_* turns the shadow rays on, 0 leaves them off.
 * Rebinding the mesh or the primitives invalidates it. */
TR_API int32_t tr_build_acceleration_structures(tr_ctx* ctx, uint64_t* address);
/* update_top_level_acceleration_structure_from_instances (src/acceleration_structures.rs:188-263, src/main.rs:1332-1345):
 * after tr_set_instances.  The handle may change; it is returned again. */
TR_API int32_t tr_update_top_level_acceleration_structure(tr_ctx* ctx, uint64_t* address);
/* trace_shadow_ray (lighting.rs:97-125) for a batch of rays given in host memory: origins/directions [n*3], t_max [n];
 * lit[i] = 1 when nothing is hit between t_min = 0.001 and t_max[i], else 0. */
TR_API int32_t tr_trace_shadow_rays(tr_ctx* ctx, uint32_t n, const float* origins, const float* directions,
                                    const float* t_max, uint8_t* lit);
/* Parity hook: the occluded-ray bits the last shading pass of `layer` used, [5][h*w] words (band rows are written):
 * planes 0-3 bit i = the ray towards the i-th light of the pixel's cluster list, plane 4 bit 0 = the sun ray. */
TR_API int32_t tr_read_shadow_mask(tr_ctx* ctx, int32_t layer, uint32_t* mask);

/* ------------------------------------------------------------------ */
/* Per-frame passes, in record() order (src/main.rs:1551-2263).        */
/* All enqueue on the context stream and return without blocking.       */
/* ------------------------------------------------------------------ */
/* frustum_culling + demultiplex_draws (shader/src/lib.rs:411-517; dispatch main.rs:1669-1838).
 * Produces instance_counts[n_prim], an ascending visible-instance list,
 * the four draw lists (ascending primitive id) and draw_counts[4]. */
TR_API int32_t tr_cull(tr_ctx* ctx, const tr_culling_push_constants* pc);
/* write_cluster_data (shader/src/lib.rs:519-594; recorded at init/resize, main.rs:1459-1517). */
TR_API int32_t tr_build_clusters(tr_ctx* ctx, const tr_write_cluster_data_push_constants* pc);
/* assign_lights_to_clusters (shader/src/lib.rs:596-645; main.rs:1766-1798); lists come out in ascending light id. */
TR_API int32_t tr_assign_lights(tr_ctx* ctx, const tr_assign_lights_push_constants* pc);
/* depth pre-passes + varyings of both layers (lib.rs:319-391; main.rs:1900-1944, 2005-2042),
 * as a software visibility pass over the culled draw lists. */
TR_API int32_t tr_visibility(tr_ctx* ctx, const tr_push_constants* pc);
/* `fragment` (lib.rs:164-249) over the opaque layer, band rows only. */
TR_API int32_t tr_shade_opaque(tr_ctx* ctx, const tr_push_constants* pc);
/* multi-GPU only: exchange opaque bands so every rank holds the whole mip-0 frame. */
TR_API int32_t tr_allgather_opaque(tr_ctx* ctx);
/* generate_mips (main.rs:2054-2063; level count main.rs:2590-2592). */
TR_API int32_t tr_generate_mips(tr_ctx* ctx);
/* `fragment_transmission` (lib.rs:37-162) over the transmissive layer, band rows only. */
TR_API int32_t tr_shade_transmission(tr_ctx* ctx, const tr_push_constants* pc);
/* fragment_tonemap (lib.rs:683-697, tonemapping.rs) -> sRGB8 RGBA. */
TR_API int32_t tr_tonemap(tr_ctx* ctx, const tr_baked_lottes_tonemapper_params* params);

/* One call = one record(): cull -> assign lights -> visibility -> opaque ->
 * [all-gather] -> mips -> transmission -> tonemap (src/main.rs:1551-2263). */
typedef struct {
    tr_culling_push_constants culling;
    tr_assign_lights_push_constants assign_lights;
    tr_push_constants push_constants;
    tr_baked_lottes_tonemapper_params tonemap;
    uint32_t flags; /* TR_FRAME_* */
} tr_frame_params;
#define TR_FRAME_SKIP_TONEMAP 1u
#define TR_FRAME_SKIP_VISIBILITY 2u /* reuse the injected / previous G-buffer */
TR_STATIC_ASSERT(sizeof(tr_frame_params) == 304, "tr_frame_params");
TR_STATIC_ASSERT(offsetof(tr_frame_params, assign_lights) == 96, "tr_frame_params.assign_lights");
TR_STATIC_ASSERT(offsetof(tr_frame_params, push_constants) == 176, "tr_frame_params.push_constants");
TR_STATIC_ASSERT(offsetof(tr_frame_params, tonemap) == 272, "tr_frame_params.tonemap");
TR_STATIC_ASSERT(offsetof(tr_frame_params, flags) == 300, "tr_frame_params.flags");
TR_API int32_t tr_frame(tr_ctx* ctx, const tr_frame_params* params);
/* Marks the start of a frame for the per-pass timers when the passes are called one by one instead of through tr_frame
 * (the reference collects its per-pass timestamp queries once per frame, src/profiling.rs:101-131). */
TR_API int32_t tr_begin_frame(tr_ctx* ctx);

/* ------------------------------------------------------------------ */
/* Parity hooks: inject / read back every intermediate                  */
/* ------------------------------------------------------------------ */
TR_API int32_t tr_set_gbuffer(tr_ctx* ctx, int32_t layer, const tr_gbuffer_planes* planes);
TR_API int32_t tr_read_gbuffer(tr_ctx* ctx, int32_t layer, const tr_gbuffer_planes_out* planes);
/* mip 0 of the sampled opaque pyramid, RGBA16F bits, full frame. */
TR_API int32_t tr_set_opaque_frame(tr_ctx* ctx, const uint16_t* rgba16f);
/* hdr_framebuffer, RGBA16F bits, full frame (e.g. to tonemap a caller-provided image). */
TR_API int32_t tr_set_hdr(tr_ctx* ctx, const uint16_t* rgba16f);
TR_API int32_t tr_set_cluster_lights(tr_ctx* ctx, const uint32_t* counts, const uint32_t* indices); /* [n_clusters], [n_clusters*128] */
TR_API int32_t tr_read_visible_instances(tr_ctx* ctx, uint32_t* ids, uint32_t capacity, uint32_t* n_visible);
TR_API int32_t tr_read_instance_counts(tr_ctx* ctx, uint32_t* counts, uint32_t capacity);
TR_API int32_t tr_read_draws(tr_ctx* ctx, uint32_t bucket, tr_draw_indexed_indirect_command* cmds,
                             uint32_t capacity, uint32_t* n_draws);
TR_API int32_t tr_read_cluster_aabbs(tr_ctx* ctx, tr_cluster_aabb* aabbs, uint32_t capacity);
TR_API int32_t tr_read_cluster_lights(tr_ctx* ctx, uint32_t* counts, uint32_t* indices); /* [n_clusters], [n_clusters*128] */
TR_API int32_t tr_read_hdr(tr_ctx* ctx, uint16_t* rgba16f);              /* hdr_framebuffer, full frame */
TR_API int32_t tr_read_hdr_f32(tr_ctx* ctx, float* rgba32f);             /* needs TR_FLAG_HDR_F32_DEBUG */
TR_API int32_t tr_read_pyramid_level(tr_ctx* ctx, uint32_t level, uint16_t* rgba16f, uint32_t* w, uint32_t* h);
TR_API int32_t tr_read_srgb8(tr_ctx* ctx, uint8_t* rgba8);
/* Same band copy, enqueued on a copy stream behind the frame just recorded, so it overlaps the next tr_frame (the
 * reference presents frame n while recording frame n+1, src/main.rs:1389-1403).  `rgba8` should be pinned and must stay
 * valid until tr_wait_readback / tr_sync returns. */
TR_API int32_t tr_read_srgb8_async(tr_ctx* ctx, uint8_t* rgba8);
TR_API int32_t tr_wait_readback(tr_ctx* ctx);
TR_API int32_t tr_mip_levels(tr_ctx* ctx, uint32_t* levels);
/* per-pass device time of the last tr_frame (profiling.rs zone taxonomy). */
typedef struct {
    float cull_ms, assign_lights_ms, visibility_ms, shade_opaque_ms, allgather_ms, mips_ms,
        shade_transmission_ms, tonemap_ms, total_ms;
} tr_frame_times;
TR_API int32_t tr_enable_timing(tr_ctx* ctx, int32_t enable);
TR_API int32_t tr_read_frame_times(tr_ctx* ctx, tr_frame_times* out);
/* per-pass device time summed over the tr_frame calls since tr_enable_timing / the previous call
 * (at most the last 128 frames are kept); the reference streams the same zones to Tracy once per
 * frame (src/profiling.rs:101-131). */
TR_API int32_t tr_read_pass_totals(tr_ctx* ctx, tr_frame_times* sum, uint32_t* n_frames);

/* ------------------------------------------------------------------ */
/* glam-pbr contract batch evaluators (device arithmetic, host buffers) */
/* ------------------------------------------------------------------ */
TR_API int32_t tr_eval_basic_brdf(tr_ctx* ctx, uint32_t n, const tr_basic_brdf_params* params, tr_brdf_result* out);
TR_API int32_t tr_eval_transmission_btdf(tr_ctx* ctx, uint32_t n, const tr_transmission_btdf_params* params, tr_vec3* out);
/* the light loop of the frame kernels, one (pixel, point light) pair per element (see tr_point_light_params) */
TR_API int32_t tr_eval_point_light(tr_ctx* ctx, uint32_t n, const tr_point_light_params* params, tr_point_light_result* out);
/* samples the context's current opaque pyramid + LUT (Appendix E rules). */
TR_API int32_t tr_eval_ibl_volume_refraction(tr_ctx* ctx, uint32_t n, const tr_mat4* proj_view,
                                             const tr_ibl_volume_refraction_params* params, tr_vec3* out);

/* ------------------------------------------------------------------ */
/* Multi-GPU: one context per rank/GPU; image bands                     */
/* ------------------------------------------------------------------ */
#define TR_NCCL_UNIQUE_ID_BYTES 128
TR_API int32_t tr_comm_unique_id(uint8_t id[TR_NCCL_UNIQUE_ID_BYTES]);
TR_API int32_t tr_comm_init(tr_ctx* ctx, const uint8_t id[TR_NCCL_UNIQUE_ID_BYTES], int32_t rank, int32_t n_ranks);
TR_API int32_t tr_comm_destroy(tr_ctx* ctx);
/* Caller-balanced bands: `bounds` = n_ranks + 1 ascending row boundaries, the same on every rank (0 ... height); rank r
 * renders rows [bounds[r], bounds[r+1]).  Equal rows (tr_comm_init's default) leave the ranks unequal work where
 * coverage and light counts vary down the frame; the frame is bitwise the same for any boundaries.  Call between
 * frames, on all ranks alike. */
TR_API int32_t tr_set_bands(tr_ctx* ctx, const uint32_t* bounds, uint32_t n_bounds);
/* Peer-store path: K4 writes its band into every peer's frame directly over NVLink. */
#define TR_IPC_HANDLE_BYTES 64
TR_API int32_t tr_peer_export(tr_ctx* ctx, uint8_t handle[TR_IPC_HANDLE_BYTES]);
TR_API int32_t tr_peer_attach(tr_ctx* ctx, int32_t rank, int32_t n_ranks, const uint8_t* handles /* [n_ranks*64] */);
/* Raw device pointers, so host frameworks (torch.distributed, a Rust host) can
 * wrap the buffers without a copy.  `what` is a TR_BUF_* value. */
enum { TR_BUF_OPAQUE_MIP0 = 0, TR_BUF_HDR = 1, TR_BUF_SRGB8 = 2, TR_BUF_HDR_F32 = 3 };
TR_API int32_t tr_device_buffer(tr_ctx* ctx, int32_t what, void** device_ptr, size_t* bytes);

/* ------------------------------------------------------------------ */
/* Diagnostics (benchmark support; not part of the reference surface)  */
/* ------------------------------------------------------------------ */
/* kernels launched by this library in this process so far */
TR_API int32_t tr_launch_count(uint64_t* out);
/* rasteriser work since the last reset: [0] box pixels binned to tiles, [1] box pixels left after hierarchical Z,
 * [2] exact (double) coverage evaluations, [3] span pixels (fp32 depth-plane tests) */
TR_API int32_t tr_raster_stats(tr_ctx* ctx, uint64_t out[4], int32_t reset);
/* measured roofline denominators on the context's GPU: dependent-FFMA chains / STREAM-style copy */
TR_API int32_t tr_measure_fp32_peak(tr_ctx* ctx, float* tflops);
TR_API int32_t tr_measure_hbm_copy(tr_ctx* ctx, float* gbs);

#ifdef __cplusplus
}
#endif
#endif /* TR_ABI_H */

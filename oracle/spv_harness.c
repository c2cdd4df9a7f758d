/* oracle/spv_harness.c — TEST INFRASTRUCTURE ONLY.
 *
 * Runs the reference's SHIPPED shader modules (compiled-shaders/normal/*.spv, translated to C by oracle/spv2c.py
 * into oracle/_ref/) on the CPU with the reference's own descriptor-set interface:
 *   frustum_culling            shader/src/lib.rs:412-440   dispatch src/main.rs:1762
 *   demultiplex_draws          shader/src/lib.rs:474-517   dispatch src/main.rs:1837
 *   write_cluster_data         shader/src/lib.rs:520-579   dispatch src/main.rs:1511-1515
 *   assign_lights_to_clusters  shader/src/lib.rs:597-645   dispatch src/main.rs:1792-1795
 *   fragment                   shader/src/lib.rs:164-249
 *   fragment_transmission      shader/src/lib.rs:37-162
 *   fragment_tonemap           shader/src/lib.rs:683-697
 *   vertex_instanced_with_scale shader/src/lib.rs:364-391
 *   vertex_instanced, depth_pre_pass_instanced, depth_pre_pass_vertex_alpha_clip   shader/src/lib.rs:335-362, 319-333, 295-317
 *   depth_pre_pass_alpha_clip  shader/src/lib.rs:270-293
 *   ray-tracing/fragment, ray-tracing/fragment_transmission   the same two entry points with trace_shadow_ray
 *                              (shader/src/lighting.rs:97-125) compiled in: SPV_KHR_ray_query
 * Together they are `oracle/_ref/libspvref.so`: outputs of the reference's own compiled code, which pin the C
 * restatement (oracle/ *.c) and, through it, the CUDA path.
 *
 * What this file supplies is what the Vulkan implementation supplied: dispatch / per-fragment invocation, the
 * varyings (decoded from the G-buffer exactly as the oracle's frame drivers do), image sampling (SURVEY.md Appendix E,
 * the oracle's orc_sample_* functions) and screen-space derivatives (the G-buffer's forward differences).
 * Invocations run one at a time in ascending id order, so atomically appended lists come out in ascending order.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "spv_ctx.h"

#define DECL(name) void spv_##name(spv_ctx*); extern const uint32_t spv_##name##_local_size[3];
DECL(frustum_culling) DECL(demultiplex_draws) DECL(write_cluster_data) DECL(assign_lights_to_clusters)
DECL(fragment) DECL(fragment_transmission) DECL(fragment_tonemap) DECL(vertex_instanced_with_scale)
DECL(vertex_instanced) DECL(depth_pre_pass_instanced) DECL(depth_pre_pass_alpha_clip) DECL(depth_pre_pass_vertex_alpha_clip)
DECL(rt_fragment) DECL(rt_fragment_transmission)   /* compiled-shaders/ray-tracing/: the same entry points built with ray queries */

#define EXPORT __attribute__((visibility("default")))

static void bind(spv_ctx* c, int set, int binding, const void* p, uint64_t size) {
    c->buf[set][binding].ptr = (uint8_t*)p;
    c->buf[set][binding].size = size;
}

static void dispatch(spv_entry_fn fn, const uint32_t ls[3], spv_ctx* c, uint32_t gx, uint32_t gy, uint32_t gz) {
    uint32_t id[3];
    c->builtin[28] = id; /* GlobalInvocationId */
    for (uint32_t z = 0; z < gz * ls[2]; z++)
        for (uint32_t y = 0; y < gy * ls[1]; y++)
            for (uint32_t x = 0; x < gx * ls[0]; x++) {
                id[0] = x; id[1] = y; id[2] = z;
                fn(c);
            }
}

static uint32_t dispatch_count(uint32_t num, uint32_t group) { return num == 0 ? 0 : (num - 1) / group + 1; } /* main.rs:2639-2641 */

/* ---- compute modules ---------------------------------------------------------------------------------------- */
EXPORT void ref_frustum_culling(const tr_instance* inst, uint32_t n_inst, const tr_primitive_info* prims, uint32_t n_prims,
                                const tr_culling_push_constants* pc, uint32_t* instance_counts) {
    spv_ctx c;
    memset(&c, 0, sizeof c);
    bind(&c, 0, 0, prims, (uint64_t)n_prims * sizeof *prims);
    bind(&c, 0, 1, instance_counts, (uint64_t)n_prims * 4);
    bind(&c, 1, 0, inst, (uint64_t)n_inst * sizeof *inst);
    c.push = (uint8_t*)pc;
    memset(instance_counts, 0, (size_t)n_prims * 4); /* "zeroing the instance count buffer", main.rs:1669-1700 */
    dispatch(spv_frustum_culling, spv_frustum_culling_local_size, &c, dispatch_count(n_inst, 64), 1, 1);
}

/* The module only counts per primitive (the reference's loader makes one instance per primitive).  Per-instance
 * visibility is read off it by dispatching one instance at a time: invocation id = instance index, everything else
 * as in the reference; visible[i] = the count its primitive received. */
EXPORT void ref_frustum_culling_bits(const tr_instance* inst, uint32_t n_inst, const tr_primitive_info* prims,
                                     uint32_t n_prims, const tr_culling_push_constants* pc, uint8_t* visible) {
    spv_ctx c;
    memset(&c, 0, sizeof c);
    uint32_t* counts = calloc(n_prims ? n_prims : 1, 4);
    bind(&c, 0, 0, prims, (uint64_t)n_prims * sizeof *prims);
    bind(&c, 0, 1, counts, (uint64_t)n_prims * 4);
    bind(&c, 1, 0, inst, (uint64_t)n_inst * sizeof *inst);
    c.push = (uint8_t*)pc;
    uint32_t id[3] = {0, 0, 0};
    c.builtin[28] = id;
    for (uint32_t i = 0; i < n_inst; i++) {
        uint32_t p = inst[i].primitive_id;
        uint32_t before = counts[p];
        id[0] = i;
        spv_frustum_culling(&c);
        visible[i] = (uint8_t)(counts[p] - before);
    }
    free(counts);
}

EXPORT void ref_demultiplex_draws(const tr_primitive_info* prims, const uint32_t* instance_counts, uint32_t n_prims,
                                  tr_draw_indexed_indirect_command* draws[4], uint32_t draw_counts[4]) {
    spv_ctx c;
    memset(&c, 0, sizeof c);
    bind(&c, 0, 0, prims, (uint64_t)n_prims * sizeof *prims);
    bind(&c, 0, 1, instance_counts, (uint64_t)n_prims * 4);
    bind(&c, 0, 2, draw_counts, 16);
    for (int k = 0; k < 4; k++) bind(&c, 0, 3 + k, draws[k], (uint64_t)n_prims * sizeof(tr_draw_indexed_indirect_command));
    memset(draw_counts, 0, 16); /* "zeroing the draw count buffer" */
    dispatch(spv_demultiplex_draws, spv_demultiplex_draws_local_size, &c, dispatch_count(n_prims, 64), 1, 1);
}

EXPORT void ref_write_cluster_data(const tr_uniforms* u, const tr_write_cluster_data_push_constants* pc, uint32_t nz,
                                   tr_cluster_aabb* out) {
    spv_ctx c;
    memset(&c, 0, sizeof c);
    uint32_t nx = u->num_clusters.x, ny = u->num_clusters.y;
    bind(&c, 0, 3, u, sizeof *u);
    bind(&c, 1, 0, out, (uint64_t)nx * ny * nz * sizeof *out);
    c.push = (uint8_t*)pc;
    dispatch(spv_write_cluster_data, spv_write_cluster_data_local_size, &c, dispatch_count(nx, 4), dispatch_count(ny, 4),
             dispatch_count(nz, 4));
}

EXPORT void ref_assign_lights_to_clusters(const tr_light* lights, uint32_t n_lights, const tr_cluster_aabb* clusters,
                                          uint32_t n_clusters, const tr_assign_lights_push_constants* pc, uint32_t* counts,
                                          uint32_t* indices) {
    spv_ctx c;
    memset(&c, 0, sizeof c);
    bind(&c, 0, 0, lights, (uint64_t)n_lights * sizeof *lights);
    bind(&c, 0, 1, counts, (uint64_t)n_clusters * 4);
    bind(&c, 0, 2, indices, (uint64_t)n_clusters * TR_MAX_LIGHTS_PER_CLUSTER * 4);
    bind(&c, 1, 0, clusters, (uint64_t)n_clusters * sizeof *clusters);
    c.push = (uint8_t*)pc;
    memset(counts, 0, (size_t)n_clusters * 4); /* cmd_fill_buffer before the dispatch, main.rs:1770-1790 */
    /* x = cluster, y = light (lib.rs:606-607); y outermost in dispatch() => every list is in ascending light id */
    dispatch(spv_assign_lights_to_clusters, spv_assign_lights_to_clusters_local_size, &c, dispatch_count(n_clusters, 8),
             dispatch_count(n_lights, 8), 1);
}

/* ---- fragment modules --------------------------------------------------------------------------------------- */
typedef struct {
    const orc_scene* s;
    const orc_pyramid* fb;
    const orc_lut* lut;
    const orc_frag_derivatives* d;
    v4 tonemap_texel;
} frag_env;

static void frag_sample(spv_ctx* c, spv_handle image, spv_handle sampler, const float* coord, int n, int has_lod, float lod,
                        float* out) {
    const frag_env* e = (const frag_env*)c->user;
    (void)n;
    out[0] = out[1] = out[2] = 0.0f;
    out[3] = 1.0f;
    if (image.set == 3) { /* framebuffer.sample_by_lod(clamp_sampler, uv, lod), lib.rs:135-138 */
        v3 r = orc_sample_pyramid(e->fb, coord[0], coord[1], has_lod ? lod : 0.0f);
        out[0] = r.x; out[1] = r.y; out[2] = r.z;
        return;
    }
    if (image.set == 1) { /* fragment_tonemap's texture at the pixel's own centre */
        out[0] = e->tonemap_texel.x; out[1] = e->tonemap_texel.y; out[2] = e->tonemap_texel.z; out[3] = e->tonemap_texel.w;
        return;
    }
    if (e->s && image.index == e->s->uniforms->ggx_lut_texture_index && sampler.binding == 4) { /* lib.rs:126-133 */
        v2 r = orc_sample_lut(e->lut, coord[0], coord[1]);
        out[0] = r.x; out[1] = r.y; out[2] = 0.0f;
        return;
    }
    if (e->s && e->s->textures && image.index < e->s->n_textures) { /* TextureSampler::sample, lib.rs:251-267 */
        v2 uv = {coord[0], coord[1]};
        v2 zero = {0.0f, 0.0f};
        v4 r = orc_sample_texture(&e->s->textures[image.index], uv, e->d ? e->d->duv_dx : zero, e->d ? e->d->duv_dy : zero);
        out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
        return;
    }
    out[0] = out[1] = out[2] = out[3] = 0.0f; /* robust access, like the oracle's tex_sample */
}

static void frag_dpd(spv_ctx* c, int is_y, int loc, int n, const float* value, float* out) {
    const frag_env* e = (const frag_env*)c->user;
    (void)value;
    for (int k = 0; k < n; k++) out[k] = 0.0f;
    if (!e->d) return;
    if (loc == 2 && n == 2) { /* uv */
        v2 d = is_y ? e->d->duv_dy : e->d->duv_dx;
        out[0] = d.x; out[1] = d.y;
    } else if (n == 3) { /* -view_vector: its differences are the world-position differences (lighting.rs:246-249) */
        v3 d = is_y ? e->d->dpos_dy : e->d->dpos_dx;
        out[0] = d.x; out[1] = d.y; out[2] = d.z;
    }
}

/* Ray queries.  The reference's loop (lighting.rs:112-118) confirms every candidate the hardware reports, so the committed
 * intersection is "the nearest triangle in [t_min, t_max], if any" whatever the geometry's opacity flag; the module only asks
 * whether there is one.  The environment therefore reports no candidate to the shader (rq_proceed = 0: the intersection is
 * committed on its own, as for opaque geometry) and answers the committed-type question with the shadow-ray definition of
 * oracle/shadow.c — which primitive/instance set is traced and how a ray meets a triangle is ours to define;
 * WHEN a ray is traced, from where, how far, and what its answer does to the light is the module's. */
static int frag_rq_proceed(spv_ctx* c, spv_rayq* q) {
    (void)c; (void)q;
    return 0;
}
static uint32_t frag_rq_committed_type(spv_ctx* c, spv_rayq* q) {
    const frag_env* e = (const frag_env*)c->user;
    /* the contract the shadow-ray definition was written to: lighting.rs:104-111 */
    if (!q->initialised || q->flags != 0u || q->cull_mask != 0xffu || q->t_min != 0.001f || !e->s->accel) abort();
    v3 o = {q->origin[0], q->origin[1], q->origin[2]}, d = {q->direction[0], q->direction[1], q->direction[2]};
    return orc_trace_shadow(e->s->accel, o, d, q->t_max) == 0.0f ? 1u : 0u;   /* 1 = committed triangle */
}

static void bind_scene(spv_ctx* c, const orc_scene* s) {
    c->push = (uint8_t*)s->pc;
    bind(c, 0, 2, s->materials, (uint64_t)s->n_materials * sizeof(tr_material_info));
    bind(c, 0, 3, s->uniforms, sizeof(tr_uniforms));
    bind(c, 2, 0, s->lights, (uint64_t)s->n_lights * sizeof(tr_light));
    bind(c, 2, 1, s->cluster_light_counts, (uint64_t)s->n_clusters * 4);
    bind(c, 2, 2, s->cluster_light_indices, (uint64_t)s->n_clusters * TR_MAX_LIGHTS_PER_CLUSTER * 4);
    c->sample = frag_sample;
    c->dpd = frag_dpd;
    c->rq_proceed = frag_rq_proceed;
    c->rq_committed_type = frag_rq_committed_type;
}

/* one invocation of `fragment`; out2 receives the second colour attachment (lib.rs:247-248) */
EXPORT v4 ref_fragment(v3 position, v3 normal, v2 uv, uint32_t material_id, v4 frag_coord, const orc_scene* s,
                       const orc_frag_derivatives* d, v4* out2) {
    spv_ctx c;
    memset(&c, 0, sizeof c);
    frag_env e = {s, NULL, NULL, d, {0, 0, 0, 0}};
    v4 o0 = {0, 0, 0, 0}, o1 = {0, 0, 0, 0};
    bind_scene(&c, s);
    c.user = &e;
    c.in_loc[0] = &position; c.in_loc[1] = &normal; c.in_loc[2] = &uv; c.in_loc[3] = &material_id;
    c.builtin[15] = &frag_coord;
    c.out_loc[0] = &o0; c.out_loc[1] = &o1;
    if (s->accel) spv_rt_fragment(&c);   /* a scene with an acceleration structure runs the ray-tracing build of the module */
    else spv_fragment(&c);
    if (out2) *out2 = o1;
    return o0;
}

EXPORT v4 ref_fragment_transmission(v3 position, v3 normal, v2 uv, uint32_t material_id, float model_scale, v4 frag_coord,
                                    const orc_scene* s, const orc_pyramid* fb, const orc_lut* lut,
                                    const orc_frag_derivatives* d) {
    spv_ctx c;
    memset(&c, 0, sizeof c);
    frag_env e = {s, fb, lut, d, {0, 0, 0, 0}};
    v4 o0 = {0, 0, 0, 0};
    bind_scene(&c, s);
    c.user = &e;
    c.in_loc[0] = &position; c.in_loc[1] = &normal; c.in_loc[2] = &uv; c.in_loc[3] = &material_id; c.in_loc[4] = &model_scale;
    c.builtin[15] = &frag_coord;
    c.out_loc[0] = &o0;
    if (s->accel) spv_rt_fragment_transmission(&c);
    else spv_fragment_transmission(&c);
    return o0;
}

/* Frame drivers with the oracle's G-buffer decode (orc_shade_frame_with, oracle/shade.c) */
typedef struct { const orc_scene* s; const orc_pyramid* fb; const orc_lut* lut; } frame_env;

static v4 cb_fragment(v3 position, v3 normal, v2 uv, uint32_t material_id, float model_scale, v4 frag_coord,
                      const orc_frag_derivatives* d, void* user) {
    const frame_env* f = (const frame_env*)user;
    (void)model_scale;
    return ref_fragment(position, normal, uv, material_id, frag_coord, f->s, d, NULL);
}
static v4 cb_fragment_transmission(v3 position, v3 normal, v2 uv, uint32_t material_id, float model_scale, v4 frag_coord,
                                   const orc_frag_derivatives* d, void* user) {
    const frame_env* f = (const frame_env*)user;
    return ref_fragment_transmission(position, normal, uv, material_id, model_scale, frag_coord, f->s, f->fb, f->lut, d);
}

EXPORT void ref_shade_opaque_frame(const orc_gbuffer* g, const orc_scene* s, uint32_t y0, uint32_t y1, float* hdr_f32,
                                   uint16_t* hdr_f16, uint16_t* opaque_f16) {
    frame_env f = {s, NULL, NULL};
    orc_shade_frame_with(g, s->pc, 0, y0, y1, cb_fragment, &f, hdr_f32, hdr_f16, opaque_f16);
}
EXPORT void ref_shade_transmission_frame(const orc_gbuffer* g, const orc_scene* s, const orc_pyramid* fb, const orc_lut* lut,
                                         uint32_t y0, uint32_t y1, float* hdr_f32, uint16_t* hdr_f16) {
    frame_env f = {s, fb, lut};
    orc_shade_frame_with(g, s->pc, 1, y0, y1, cb_fragment_transmission, &f, hdr_f32, hdr_f16, NULL);
}

/* fragment_tonemap over pixels: the sampled texel is the pixel's own (uv at the pixel centre); returns linear rgb */
EXPORT void ref_tonemap_pixels(uint32_t n, const float* rgba_in, const tr_baked_lottes_tonemapper_params* p, float* rgba_out) {
    for (uint32_t i = 0; i < n; i++) {
        spv_ctx c;
        memset(&c, 0, sizeof c);
        frag_env e = {NULL, NULL, NULL, NULL, {rgba_in[i * 4], rgba_in[i * 4 + 1], rgba_in[i * 4 + 2], rgba_in[i * 4 + 3]}};
        v2 uv = {0.5f, 0.5f};
        c.user = &e;
        c.sample = frag_sample;
        c.push = (uint8_t*)p;
        c.in_loc[0] = &uv;
        c.out_loc[0] = rgba_out + (size_t)i * 4;
        spv_fragment_tonemap(&c);
    }
}

/* vertex_instanced_with_scale for a batch of (vertex, instance) pairs */
EXPORT void ref_vertex_instanced_with_scale(uint32_t n, const float* positions, const float* normals, const float* uvs,
                                            const uint32_t* instance_index, const tr_instance* inst, uint32_t n_inst,
                                            const tr_push_constants* pc, float* clip /*[n*4]*/, float* out_position /*[n*3]*/,
                                            float* out_normal /*[n*3]*/, float* out_uv /*[n*2]*/, uint32_t* out_material,
                                            float* out_scale) {
    for (uint32_t i = 0; i < n; i++) {
        spv_ctx c;
        memset(&c, 0, sizeof c);
        bind(&c, 1, 0, inst, (uint64_t)n_inst * sizeof *inst);
        c.push = (uint8_t*)pc;
        int32_t ii = (int32_t)instance_index[i];
        c.in_loc[0] = (void*)(positions + (size_t)i * 3);
        c.in_loc[1] = (void*)(normals + (size_t)i * 3);
        c.in_loc[2] = (void*)(uvs + (size_t)i * 2);
        c.builtin[43] = &ii;
        c.builtin[0] = clip + (size_t)i * 4;
        c.out_loc[0] = out_position + (size_t)i * 3;
        c.out_loc[1] = out_normal + (size_t)i * 3;
        c.out_loc[2] = out_uv + (size_t)i * 2;
        c.out_loc[3] = out_material + i;
        c.out_loc[4] = out_scale + i;
        spv_vertex_instanced_with_scale(&c);
    }
}

/* The other three vertex stages of the path for a batch of (vertex, instance) pairs; outputs a stage does not have are left
 * untouched.  which = 1 vertex_instanced (lib.rs:335-362: the opaque G-buffer pass), 2 depth_pre_pass_instanced (:319-333: the
 * clip position the visibility pass rasterises), 3 depth_pre_pass_vertex_alpha_clip (:295-317: clip, uv, material id). */
EXPORT int ref_vertex_stage(uint32_t which, uint32_t n, const float* positions, const float* normals, const float* uvs,
                            const uint32_t* instance_index, const tr_instance* inst, uint32_t n_inst, const tr_push_constants* pc,
                            float* clip /*[n*4]*/, float* out_position /*[n*3]*/, float* out_normal /*[n*3]*/, float* out_uv /*[n*2]*/,
                            uint32_t* out_material) {
    if (which < 1 || which > 3) return -1;
    for (uint32_t i = 0; i < n; i++) {
        spv_ctx c;
        memset(&c, 0, sizeof c);
        bind(&c, 1, 0, inst, (uint64_t)n_inst * sizeof *inst);
        c.push = (uint8_t*)pc;
        int32_t ii = (int32_t)instance_index[i];
        c.builtin[43] = &ii;
        c.builtin[0] = clip + (size_t)i * 4;
        c.in_loc[0] = (void*)(positions + (size_t)i * 3);
        if (which == 1) {
            c.in_loc[1] = (void*)(normals + (size_t)i * 3);
            c.in_loc[2] = (void*)(uvs + (size_t)i * 2);
            c.out_loc[0] = out_position + (size_t)i * 3;
            c.out_loc[1] = out_normal + (size_t)i * 3;
            c.out_loc[2] = out_uv + (size_t)i * 2;
            c.out_loc[3] = out_material + i;
            spv_vertex_instanced(&c);
        } else if (which == 2) {
            spv_depth_pre_pass_instanced(&c);
        } else {
            c.in_loc[1] = (void*)(uvs + (size_t)i * 2);
            c.out_loc[0] = out_uv + (size_t)i * 2;
            c.out_loc[1] = out_material + i;
            spv_depth_pre_pass_vertex_alpha_clip(&c);
        }
    }
    return 0;
}

/* depth_pre_pass_alpha_clip: 1 = the fragment was discarded */
EXPORT void ref_alpha_clip(uint32_t n, const float* uvs, const float* duv /*[n*4] or NULL*/, const uint32_t* material_id,
                           const orc_scene* s, uint8_t* killed) {
    for (uint32_t i = 0; i < n; i++) {
        spv_ctx c;
        memset(&c, 0, sizeof c);
        orc_frag_derivatives d;
        memset(&d, 0, sizeof d);
        if (duv) {
            d.duv_dx.x = duv[i * 4]; d.duv_dx.y = duv[i * 4 + 1];
            d.duv_dy.x = duv[i * 4 + 2]; d.duv_dy.y = duv[i * 4 + 3];
        }
        frag_env e = {s, NULL, NULL, &d, {0, 0, 0, 0}};
        c.user = &e;
        c.sample = frag_sample;
        c.dpd = frag_dpd;
        bind(&c, 0, 2, s->materials, (uint64_t)s->n_materials * sizeof(tr_material_info));
        int32_t mid = (int32_t)material_id[i];
        c.in_loc[0] = (void*)(uvs + (size_t)i * 2);
        c.in_loc[1] = &mid;
        spv_depth_pre_pass_alpha_clip(&c);
        killed[i] = (uint8_t)c.killed;
    }
}

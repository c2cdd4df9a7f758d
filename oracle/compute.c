/* oracle/compute.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of the reference's compute entry points:
 *   shader/src/lib.rs:411-469  frustum_culling + cull
 *   shader/src/lib.rs:471-517  demultiplex_draws
 *   shader/src/lib.rs:519-594  write_cluster_data + line_intersection_to_z_plane
 *   shader/src/lib.rs:596-645  assign_lights_to_clusters
 *   shared-structs/src/lib.rs:65-67, 235-241, 290-320
 * The reference appends to its lists in atomic-race order
 * (shader/src/asm.rs:3-31); the oracle (and the CUDA kernels) emit the same
 * SETS in ascending id order, which makes the per-pixel light sums
 * deterministic.  Transcendentals used in discrete decisions
 * (powf in slice_to_depth, cos/sin in cull_spotlight) are evaluated in double
 * and rounded to f32 on both sides (DESIGN.md "discrete decisions").
 * PARITY: pinned bit-exact to the reference's compiled frustum_culling.spv, demultiplex_draws.spv, write_cluster_data.spv and
 * assign_lights_to_clusters.spv (oracle.h, tests/test_reference_spirv.py).
 */
#include "oracle.h"

static v3 xyz4(tr_vec4 a) { return v3_new(a.x, a.y, a.z); }
static v3 xyz3a(tr_vec3a a) { return v3_new(a.x, a.y, a.z); }
static void load_m4(const tr_mat4* in, m4* out) { memcpy(out, in, sizeof(m4)); }

/* shader/src/lib.rs:442-469; Similarity * Vec3 = shared-structs/src/lib.rs:235-241 */
int orc_cull(tr_vec4 sphere, const tr_packed_similarity* t, const tr_culling_push_constants* pc) {
    v3 translation = xyz4(t->translation_and_scale);
    float scale = t->translation_and_scale.w;
    v4 rot = v4_new(t->rotation.x, t->rotation.y, t->rotation.z, t->rotation.w);

    v3 center = xyz4(sphere);
    center = v3_add(translation, v3_scale(quat_mul_v3(rot, center), scale));
    m4 view;
    load_m4(&pc->view, &view);
    v4 c4 = m4_mul_v4(&view, v4_new(center.x, center.y, center.z, 1.0f));
    center = v3_new(c4.x, c4.y, -c4.z); /* :452 */

    float radius = sphere.w;
    radius *= scale;

    int visible = center.z + radius > pc->z_near;
    visible &= center.z * pc->frustum_x_xz.y - fabsf(center.x) * pc->frustum_x_xz.x < radius;
    visible &= center.z * pc->frustum_y_yz.y - fabsf(center.y) * pc->frustum_y_yz.x < radius;
    return !visible;
}

/* shader/src/lib.rs:411-440 */
void orc_frustum_culling(const tr_instance* inst, uint32_t n_inst, const tr_primitive_info* prims, uint32_t n_prims,
                         const tr_culling_push_constants* pc, uint32_t* instance_counts, uint32_t* visible_ids,
                         uint32_t* n_visible) {
    for (uint32_t p = 0; p < n_prims; p++) instance_counts[p] = 0;
    uint32_t nv = 0;
    for (uint32_t i = 0; i < n_inst; i++) {
        const tr_primitive_info* prim = &prims[inst[i].primitive_id];
        if (orc_cull(prim->packed_bounding_sphere, &inst[i].transform, pc)) continue;
        instance_counts[inst[i].primitive_id]++;
        if (visible_ids) visible_ids[nv] = i;
        nv++;
    }
    if (n_visible) *n_visible = nv;
}

/* shader/src/lib.rs:471-517 */
void orc_demultiplex_draws(const tr_primitive_info* prims, const uint32_t* instance_counts, uint32_t n_prims,
                           tr_draw_indexed_indirect_command* draws[4], uint32_t draw_counts[4]) {
    for (int b = 0; b < 4; b++) draw_counts[b] = 0;
    for (uint32_t d = 0; d < n_prims; d++) {
        uint32_t count = instance_counts[d];
        if (count == 0) continue;
        const tr_primitive_info* prim = &prims[d];
        uint32_t bucket = prim->draw_buffer_index < 3 ? prim->draw_buffer_index : 3; /* match arm `_` */
        uint32_t slot = draw_counts[bucket]++;
        tr_draw_indexed_indirect_command c;
        c.instance_count = count;
        c.index_count = prim->index_count;
        c.first_index = prim->first_index;
        c.first_instance = prim->first_instance;
        c.vertex_offset = 0;
        draws[bucket][slot] = c;
    }
}

/* shared-structs/src/lib.rs:65-67 */
float orc_slice_to_depth(const tr_light_cluster_coefficients* c, uint32_t slice) {
    float ratio = c->z_far / c->z_near;
    float t = (float)slice / (float)c->num_depth_slices;
    return -c->z_near * (float)pow((double)ratio, (double)t);
}

/* shader/src/lib.rs:583-594 */
static v3 line_intersection_to_z_plane(v3 a, v3 b, float z_distance) {
    v3 normal = v3_new(0.0f, 0.0f, 1.0f);
    v3 a_to_b = v3_sub(b, a);
    float t = (z_distance - v3_dot(normal, a)) / v3_dot(normal, a_to_b);
    return v3_add(a, v3_scale(a_to_b, t));
}

/* shader/src/lib.rs:519-580 */
void orc_write_cluster_data(const tr_uniforms* u, const tr_write_cluster_data_push_constants* pc, uint32_t nz,
                            tr_cluster_aabb* out) {
    m4 inv_persp;
    load_m4(&pc->inverse_perspective, &inv_persp);
    uint32_t nx = u->num_clusters.x, ny = u->num_clusters.y;
    for (uint32_t iz = 0; iz < nz; iz++)
        for (uint32_t iy = 0; iy < ny; iy++)
            for (uint32_t ix = 0; ix < nx; ix++) {
                uint32_t cluster_id = iz * nx * ny + iy * nx + ix;
                float smin[2] = {(float)ix * u->cluster_size_in_pixels.x, (float)iy * u->cluster_size_in_pixels.y};
                float smax[2] = {(float)(ix + 1) * u->cluster_size_in_pixels.x,
                                 (float)(iy + 1) * u->cluster_size_in_pixels.y};
                v3 view_space[2];
                for (int k = 0; k < 2; k++) {
                    const float* sp = k == 0 ? smin : smax;
                    /* screen_to_clip :540-544 */
                    float px = sp[0] / (float)pc->screen_dimensions.x;
                    float py = sp[1] / (float)pc->screen_dimensions.y;
                    px = px * 2.0f - 1.0f;
                    py = py * 2.0f - 1.0f;
                    /* clip_to_view :546-550 */
                    v4 v = m4_mul_v4(&inv_persp, v4_new(px, py, 0.0f, 1.0f));
                    view_space[k] = v3_divs(v3_new(v.x, v.y, v.z), v.w);
                }
                float z_near = orc_slice_to_depth(&u->light_clustering_coefficients, iz);
                float z_far = orc_slice_to_depth(&u->light_clustering_coefficients, iz + 1);
                v3 eye = v3_new(0.0f, 0.0f, 1.0f); /* :560 (sic) */
                v3 min_near = line_intersection_to_z_plane(eye, view_space[0], z_near);
                v3 min_far = line_intersection_to_z_plane(eye, view_space[0], z_far);
                v3 max_near = line_intersection_to_z_plane(eye, view_space[1], z_near);
                v3 max_far = line_intersection_to_z_plane(eye, view_space[1], z_far);
                v3 lo = v3_min(v3_min(v3_min(min_near, min_far), max_near), max_far);
                v3 hi = v3_max(v3_max(v3_max(min_near, min_far), max_near), max_far);
                tr_cluster_aabb c;
                c.min.x = lo.x; c.min.y = lo.y; c.min.z = lo.z; c.min._pad = 0.0f;
                c.max.x = hi.x; c.max.y = hi.y; c.max.z = hi.z; c.max._pad = 0.0f;
                out[cluster_id] = c;
            }
}

/* shared-structs/src/lib.rs:291-298 */
float orc_aabb_distance_sq(const tr_cluster_aabb* a, v3 p) {
    v3 d = v3_max(v3_max(v3_sub(xyz3a(a->min), p), v3_sub(p, xyz3a(a->max))), v3_splat(0.0f));
    return v3_length_squared(d);
}

/* shared-structs/src/lib.rs:301-319 */
int orc_aabb_cull_spotlight(const tr_cluster_aabb* a, v3 origin, v3 direction, float angle, float range) {
    v3 center = v3_divs(v3_add(xyz3a(a->min), xyz3a(a->max)), 2.0f);
    float radius = v3_length(v3_sub(xyz3a(a->max), center));
    v3 vector = v3_sub(center, origin);
    float vector_len_sq = v3_dot(vector, vector);
    float vector_1_len = v3_dot(vector, direction);
    float vector_1_len_sq = vector_1_len * vector_1_len;
    float ca = (float)cos((double)angle), sa = (float)sin((double)angle);
    float distance_closest_point = ca * sqrtf(vector_len_sq - vector_1_len_sq) - vector_1_len * sa;
    int angle_cull = distance_closest_point > radius;
    int front_cull = vector_1_len > radius + range;
    int back_cull = vector_1_len < -radius;
    return angle_cull || front_cull || back_cull;
}

/* shader/src/lib.rs:596-645, lists in ascending light id */
void orc_assign_lights_to_clusters(const tr_light* lights, uint32_t n_lights, const tr_cluster_aabb* clusters,
                                   uint32_t n_clusters, const tr_assign_lights_push_constants* pc, uint32_t* counts,
                                   uint32_t* indices) {
    m4 view;
    load_m4(&pc->view_matrix, &view);
    v4 rot = v4_new(pc->view_rotation.x, pc->view_rotation.y, pc->view_rotation.z, pc->view_rotation.w);
#pragma omp parallel for schedule(static)
    for (int64_t cc = 0; cc < (int64_t)n_clusters; cc++) {
        uint32_t c = (uint32_t)cc;
        uint32_t n = 0;
        for (uint32_t l = 0; l < n_lights; l++) {
            const tr_light* light = &lights[l];
            v4 lp = m4_mul_v4(&view, v4_new(light->position_and_spotlight_epsilon.x,
                                            light->position_and_spotlight_epsilon.y,
                                            light->position_and_spotlight_epsilon.z, 1.0f));
            v3 light_position = v3_new(lp.x, lp.y, lp.z);
            float falloff_distance_sq = light->colour_emission_and_falloff_distance_sq.w;
            if (orc_aabb_distance_sq(&clusters[c], light_position) > falloff_distance_sq) continue;
            if (light->spotlight_direction_and_outer_angle.w != 0.0f) {
                v3 dir = quat_mul_v3(rot, xyz4(light->spotlight_direction_and_outer_angle));
                float angle = light->spotlight_direction_and_outer_angle.w;
                float range = light->colour_emission_and_falloff_distance_sq.w;
                if (orc_aabb_cull_spotlight(&clusters[c], light_position, dir, angle, range)) continue;
            }
            /* a 129th light would overflow into the next cluster's slots in the
             * reference (lib.rs:640-644, no bound check); we drop it and saturate. */
            if (n < TR_MAX_LIGHTS_PER_CLUSTER) indices[(size_t)c * TR_MAX_LIGHTS_PER_CLUSTER + n] = l;
            if (n < TR_MAX_LIGHTS_PER_CLUSTER) n++;
        }
        counts[c] = n;
    }
}

/* inverse(proj_view) for the G-buffer position decode: cofactor expansion in
 * double (products of two f32 are exact in double), rounded to f32 once. */
void orc_mat4_inverse(const tr_mat4* m, tr_mat4* out) {
    double a[16], inv[16];
    const float* f = (const float*)m;
    for (int i = 0; i < 16; i++) a[i] = (double)f[i];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    double r = 1.0 / det;
    float* o = (float*)out;
    for (int i = 0; i < 16; i++) o[i] = (float)(inv[i] * r);
}

/* oracle/raster.c — TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Software stand-in for the fixed-function part of the reference's depth
 * pre-passes and varying interpolation:
 *   shader/src/lib.rs:319-333  depth_pre_pass_instanced   (clip = proj_view * (T * pos))
 *   shader/src/lib.rs:335-391  vertex_instanced[_with_scale] (rotation * normal, uv, material_id, scale)
 *   src/pipelines.rs:311,350-371  back-face cull, depth GREATER + write, reversed-Z
 *   src/main.rs:1586-1591  depth clear 0.0;  src/main.rs:1900-1944, 2005-2042 pass order
 * The rasteriser itself was Vulkan hardware in the reference, so the rules
 * below are OURS (DESIGN.md "visibility"): homogeneous edge functions in
 * double (no clipping needed for triangles that cross the camera plane),
 * pixel centres at +0.5, front face = counter-clockwise seen from outside
 * (glTF), a top-left style tie rule that is exact for shared edges, nearest
 * fragment wins with ties broken by the smaller global triangle id, and the
 * transmissive layer is depth-tested GREATER against the final opaque depth.
 * PARITY: NOT PINNABLE — the reference rasterises in fixed-function hardware; the rules are ours (oracle.h), and
 * self-consistency with the CUDA kernel is checked bit for bit.  The vertex stage that feeds it IS pinned
 * (vertex_instanced_with_scale.spv, byte-equal).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>

typedef struct {
    double A[3], B[3], C[3]; /* edge i is opposite (reordered) vertex i */
    float Z[3], W[3];
    uint32_t vid[3];         /* mesh vertex ids in reordered order */
    int x_lo, x_hi, y_lo, y_hi;
} tri_setup;

static v3 sim_mul(const tr_packed_similarity* t, v3 p) {
    v3 translation = v3_new(t->translation_and_scale.x, t->translation_and_scale.y, t->translation_and_scale.z);
    v4 rot = v4_new(t->rotation.x, t->rotation.y, t->rotation.z, t->rotation.w);
    return v3_add(translation, v3_scale(quat_mul_v3(rot, p), t->translation_and_scale.w));
}

/* vertex_instanced_with_scale, shader/src/lib.rs:364-391 (vertex_instanced :336-362 is the same without the scale) */
void orc_vertex_instanced_with_scale(v3 position, v3 normal, const tr_instance* instance, const tr_push_constants* pc,
                                     v4* clip, v3* out_position, v3* out_normal, uint32_t* out_material_id, float* out_scale) {
    m4 pv;
    memcpy(&pv, &pc->proj_view, sizeof pv);
    v3 wp = sim_mul(&instance->transform, position);
    v4 rot = v4_new(instance->transform.rotation.x, instance->transform.rotation.y, instance->transform.rotation.z,
                    instance->transform.rotation.w);
    *out_position = wp;
    *out_normal = quat_mul_v3(rot, normal);
    *out_material_id = instance->material_id;
    *out_scale = instance->transform.translation_and_scale.w;
    *clip = m4_mul_v4(&pv, v4_new(wp.x, wp.y, wp.z, 1.0f));
}

/* depth_pre_pass_alpha_clip, shader/src/lib.rs:269-293: 1 = the fragment is discarded */
int orc_alpha_clip_kills(const tr_material_info* m, const orc_texture* textures, uint32_t n_textures, v2 uv, v2 duv_dx,
                         v2 duv_dy) {
    float alpha = m->diffuse_factor.w;
    if (m->textures.diffuse != -1) {
        float a = 0.0f;
        if (textures && (uint32_t)m->textures.diffuse < n_textures)
            a = orc_sample_texture(&textures[m->textures.diffuse], uv, duv_dx, duv_dy).w;
        alpha *= a;
    }
    return alpha < m->alpha_clipping_cutoff;
}

/* returns 0 if the triangle is culled / off-band */
static int setup_triangle(const orc_mesh* mesh, const tr_instance* inst, const tr_primitive_info* prim, uint32_t tri,
                          const m4* pv, uint32_t width, uint32_t height, uint32_t y0, uint32_t y1, tri_setup* s) {
    uint32_t vid[3];
    float sx[3], sy[3], Z[3], W[3];
    float half_w = (float)width * 0.5f, half_h = (float)height * 0.5f;
    for (int k = 0; k < 3; k++) {
        vid[k] = mesh->indices[prim->first_index + tri * 3 + k];
        v3 p = v3_new(mesh->positions[vid[k] * 3], mesh->positions[vid[k] * 3 + 1], mesh->positions[vid[k] * 3 + 2]);
        v3 wp = sim_mul(&inst->transform, p);
        v4 c = m4_mul_v4(pv, v4_new(wp.x, wp.y, wp.z, 1.0f));
        if (!(isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && isfinite(c.w))) return 0;
        sx[k] = (c.x + c.w) * half_w;
        sy[k] = (c.y + c.w) * half_h;
        Z[k] = c.z;
        W[k] = c.w;
    }
    /* facing: det of [[sx,sy,w]] rows; glTF CCW front faces give det < 0 in y-down pixel space */
    double a0 = (double)sy[1] * W[2] - (double)W[1] * sy[2];
    double b0 = (double)W[1] * sx[2] - (double)sx[1] * W[2];
    double c0 = (double)sx[1] * sy[2] - (double)sy[1] * sx[2];
    double det = ((double)sx[0] * a0 + (double)sy[0] * b0) + (double)W[0] * c0;
    if (!(det < 0.0)) return 0;
    /* reorder (v0, v2, v1) so that the interior has positive edge functions */
    const int order[3] = {0, 2, 1};
    float rx[3], ry[3];
    for (int k = 0; k < 3; k++) {
        rx[k] = sx[order[k]];
        ry[k] = sy[order[k]];
        s->Z[k] = Z[order[k]];
        s->W[k] = W[order[k]];
        s->vid[k] = vid[order[k]];
    }
    for (int i = 0; i < 3; i++) {
        int a = (i + 1) % 3, b = (i + 2) % 3;
        s->A[i] = (double)ry[a] * s->W[b] - (double)s->W[a] * ry[b];
        s->B[i] = (double)s->W[a] * rx[b] - (double)rx[a] * s->W[b];
        s->C[i] = (double)rx[a] * ry[b] - (double)ry[a] * rx[b];
    }
    /* Vulkan clip volume: w > 0 and z <= w.  A triangle with every vertex at/behind the camera plane has no
     * fragment (eval_pixel requires w > 0). */
    if (!(s->W[0] > 0.0f) && !(s->W[1] > 0.0f) && !(s->W[2] > 0.0f)) return 0;
    /* bounding box (conservative) */
    int x_lo = 0, x_hi = (int)width - 1, y_lo = (int)y0, y_hi = (int)y1 - 1;
    float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY, pad = 0.0f;
    int whole = 0;
    if (s->W[0] > 0.0f && s->W[1] > 0.0f && s->W[2] > 0.0f) {
        for (int k = 0; k < 3; k++) {
            float px = rx[k] / s->W[k], py = ry[k] / s->W[k];
            mnx = f_min(mnx, px); mxx = f_max(mxx, px);
            mny = f_min(mny, py); mxy = f_max(mxy, py);
        }
    } else {
        /* the triangle crosses the camera plane: bound the part inside the near plane (z <= w), whose corners
         * are the kept vertices and the edge/near-plane intersections; one pixel of padding */
        float nd[3];
        int inside = 0;
        for (int k = 0; k < 3; k++) {
            nd[k] = s->W[k] - s->Z[k];
            if (nd[k] >= 0.0f) inside++;
        }
        if (inside == 0) return 0;
        pad = 1.0f;
        for (int k = 0; k < 3; k++) {
            int j = (k + 1) % 3;
            if (nd[k] >= 0.0f) {
                if (!(s->W[k] > 0.0f)) whole = 1;
                else {
                    float px = rx[k] / s->W[k], py = ry[k] / s->W[k];
                    mnx = f_min(mnx, px); mxx = f_max(mxx, px);
                    mny = f_min(mny, py); mxy = f_max(mxy, py);
                }
            }
            if ((nd[k] >= 0.0f) != (nd[j] >= 0.0f)) {
                int a = nd[k] >= 0.0f ? k : j, b = nd[k] >= 0.0f ? j : k; /* a inside, b outside */
                float t = nd[a] / (nd[a] - nd[b]);
                float cx = rx[a] + t * (rx[b] - rx[a]);
                float cy = ry[a] + t * (ry[b] - ry[a]);
                float cw = s->W[a] + t * (s->W[b] - s->W[a]);
                if (!(cw > 0.0f)) whole = 1;
                else {
                    float px = cx / cw, py = cy / cw;
                    mnx = f_min(mnx, px); mxx = f_max(mxx, px);
                    mny = f_min(mny, py); mxy = f_max(mxy, py);
                }
            }
        }
    }
    if (!whole) {
        mnx -= pad; mny -= pad; mxx += pad; mxy += pad;
        if (!(mxx >= 0.0f) || !(mnx <= (float)width) || !(mxy >= (float)y0) || !(mny <= (float)y1)) return 0;
        float fx_lo = floorf(mnx - 0.5f), fx_hi = ceilf(mxx - 0.5f);
        float fy_lo = floorf(mny - 0.5f), fy_hi = ceilf(mxy - 0.5f);
        if (fx_lo > (float)x_lo) x_lo = (int)fx_lo;
        if (fx_hi < (float)x_hi) x_hi = (int)fx_hi;
        if (fy_lo > (float)y_lo) y_lo = (int)fy_lo;
        if (fy_hi < (float)y_hi) y_hi = (int)fy_hi;
    }
    if (x_lo > x_hi || y_lo > y_hi) return 0;
    s->x_lo = x_lo; s->x_hi = x_hi; s->y_lo = y_lo; s->y_hi = y_hi;
    return 1;
}

/* coverage + perspective-correct barycentrics at pixel centre; returns 0 if outside */
static int eval_pixel(const tri_setup* s, int px, int py, float l[3], float* depth) {
    double qx = (double)px + 0.5, qy = (double)py + 0.5;
    double E[3];
    for (int i = 0; i < 3; i++) {
        E[i] = (s->A[i] * qx + s->B[i] * qy) + s->C[i];
        if (E[i] < 0.0) return 0;
        if (E[i] == 0.0 && !(s->A[i] > 0.0 || (s->A[i] == 0.0 && s->B[i] > 0.0))) return 0;
    }
    double S = (E[0] + E[1]) + E[2];
    if (!(S > 0.0)) return 0;
    double r = 1.0 / S; /* one reciprocal, three products: the rule the CUDA kernel evaluates too */
    l[0] = (float)(E[0] * r);
    l[1] = (float)(E[1] * r);
    l[2] = (float)(E[2] * r);
    float zq = (l[0] * s->Z[0] + l[1] * s->Z[1]) + l[2] * s->Z[2];
    float wq = (l[0] * s->W[0] + l[1] * s->W[1]) + l[2] * s->W[2];
    if (!(wq > 0.0f)) return 0;             /* Vulkan clip volume: w > 0 ... */
    float d = zq / wq;
    if (!(d > 0.0f) || d > 1.0f) return 0; /* ... and 0 <= z <= w; depth 0 is "empty" */
    *depth = d;
    return 1;
}

/* barycentrics / depth of the triangle's plane at a pixel centre WITHOUT the coverage test: what a helper invocation of
 * the 2x2 quad evaluates for the derivatives.  Returns 0 where the plane has no valid perspective division. */
static int eval_plane(const tri_setup* s, int px, int py, float l[3], float* depth) {
    double qx = (double)px + 0.5, qy = (double)py + 0.5;
    double E[3];
    for (int i = 0; i < 3; i++) E[i] = (s->A[i] * qx + s->B[i] * qy) + s->C[i];
    double S = (E[0] + E[1]) + E[2];
    if (!(S > 0.0)) return 0;
    double r = 1.0 / S;
    l[0] = (float)(E[0] * r);
    l[1] = (float)(E[1] * r);
    l[2] = (float)(E[2] * r);
    float zq = (l[0] * s->Z[0] + l[1] * s->Z[1]) + l[2] * s->Z[2];
    float wq = (l[0] * s->W[0] + l[1] * s->W[1]) + l[2] * s->W[2];
    if (!(wq > 0.0f)) return 0;
    *depth = zq / wq;
    return 1;
}

static void atomic_max_u64(uint64_t* addr, uint64_t v) {
    uint64_t cur = __atomic_load_n(addr, __ATOMIC_RELAXED);
    while (cur < v && !__atomic_compare_exchange_n(addr, &cur, v, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
}

static uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

void orc_visibility(const orc_mesh* mesh, const tr_instance* inst, uint32_t n_inst, const tr_primitive_info* prims,
                    uint32_t n_prims, const uint32_t* visible_ids, uint32_t n_visible, const tr_push_constants* pc,
                    uint32_t y0, uint32_t y1, float* depth0, float* normal0, float* uv0, uint32_t* mat0, float* depth1,
                    float* normal1, float* uv1, uint32_t* mat1, float* scale1, float* duv0, float* ddepth0, float* duv1,
                    float* ddepth1, const tr_material_info* materials, const orc_texture* textures, uint32_t n_textures) {
    (void)n_prims;
    uint32_t width = pc->framebuffer_size.x, height = pc->framebuffer_size.y;
    m4 pv;
    memcpy(&pv, &pc->proj_view, sizeof(m4));
    size_t npx = (size_t)width * height;
    uint64_t* vis[2];
    vis[0] = (uint64_t*)calloc(npx, 8);
    vis[1] = (uint64_t*)calloc(npx, 8);
    /* global triangle ids: exclusive prefix over ALL instances, by instance id */
    uint32_t* tri_prefix = (uint32_t*)malloc(((size_t)n_inst + 1) * 4);
    tri_prefix[0] = 0;
    for (uint32_t i = 0; i < n_inst; i++) tri_prefix[i + 1] = tri_prefix[i] + prims[inst[i].primitive_id].index_count / 3;

    for (int layer = 0; layer < 2; layer++) {
        /* draw buffers 0 (opaque) and 1 (alpha clip) share the opaque layer, 2 and 3 the transmissive one
         * (src/main.rs:1900-1944, 2005-2042) */
#pragma omp parallel for schedule(dynamic, 8)
        for (int64_t vv = 0; vv < (int64_t)n_visible; vv++) {
            uint32_t ii = visible_ids[vv];
            const tr_primitive_info* prim = &prims[inst[ii].primitive_id];
            if (prim->draw_buffer_index > 3u || (int)(prim->draw_buffer_index >> 1) != layer) continue;
            int alpha_clip = (prim->draw_buffer_index & 1u) != 0;
            uint32_t ntri = prim->index_count / 3;
            for (uint32_t t = 0; t < ntri; t++) {
                tri_setup s;
                if (!setup_triangle(mesh, &inst[ii], prim, t, &pv, width, height, y0, y1, &s)) continue;
                uint32_t gtid = tri_prefix[ii] + t;
                for (int py = s.y_lo; py <= s.y_hi; py++)
                    for (int px = s.x_lo; px <= s.x_hi; px++) {
                        float l[3], d;
                        if (!eval_pixel(&s, px, py, l, &d)) continue;
                        if (alpha_clip && materials) { /* depth_pre_pass_alpha_clip, shader/src/lib.rs:269-293 */
                            const tr_material_info* m = &materials[inst[ii].material_id];
                            if (m->textures.diffuse != -1) {
                                const float *u0 = &mesh->uvs[s.vid[0] * 2], *u1 = &mesh->uvs[s.vid[1] * 2], *u2 = &mesh->uvs[s.vid[2] * 2];
                                v2 uv = {(l[0] * u0[0] + l[1] * u1[0]) + l[2] * u2[0], (l[0] * u0[1] + l[1] * u1[1]) + l[2] * u2[1]};
                                v2 dq[2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};
                                for (int k = 0; k < 2; k++) {
                                    float ln[3], dn;
                                    if (eval_plane(&s, px + (k == 0), py + (k == 1), ln, &dn)) {
                                        dq[k].x = ((ln[0] * u0[0] + ln[1] * u1[0]) + ln[2] * u2[0]) - uv.x;
                                        dq[k].y = ((ln[0] * u0[1] + ln[1] * u1[1]) + ln[2] * u2[1]) - uv.y;
                                    }
                                }
                                if (orc_alpha_clip_kills(m, textures, n_textures, uv, dq[0], dq[1])) continue;
                            } else {
                                v2 zero = {0.0f, 0.0f};
                                if (orc_alpha_clip_kills(m, textures, n_textures, zero, zero, zero)) continue;
                            }
                        }
                        size_t i = (size_t)py * width + px;
                        if (layer == 1) { /* GREATER against the opaque depth already in the shared depth buffer */
                            float dop;
                            uint32_t ob = (uint32_t)(vis[0][i] >> 32);
                            memcpy(&dop, &ob, 4);
                            if (!(d > dop)) continue;
                        }
                        uint64_t key = ((uint64_t)f32_bits(d) << 32) | (uint64_t)(0xffffffffu - gtid);
                        atomic_max_u64(&vis[layer][i], key);
                    }
            }
        }
    }

    /* resolve: recompute the winning triangle at each pixel and interpolate its varyings */
    for (int layer = 0; layer < 2; layer++) {
        float* depth = layer == 0 ? depth0 : depth1;
        float* normal = layer == 0 ? normal0 : normal1;
        float* uv = layer == 0 ? uv0 : uv1;
        uint32_t* mat = layer == 0 ? mat0 : mat1;
        float* duv = layer == 0 ? duv0 : duv1;
        float* ddepth = layer == 0 ? ddepth0 : ddepth1;
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
            for (uint32_t x = 0; x < width; x++) {
                size_t i = (size_t)yy * width + x;
                uint64_t key = vis[layer][i];
                if (key == 0) {
                    depth[i] = 0.0f;
                    normal[i * 3] = normal[i * 3 + 1] = normal[i * 3 + 2] = 0.0f;
                    uv[i * 2] = uv[i * 2 + 1] = 0.0f;
                    mat[i] = 0xffffffffu;
                    if (layer == 1 && scale1) scale1[i] = 0.0f;
                    if (duv) duv[i * 4] = duv[i * 4 + 1] = duv[i * 4 + 2] = duv[i * 4 + 3] = 0.0f;
                    if (ddepth) ddepth[i * 2] = ddepth[i * 2 + 1] = 0.0f;
                    continue;
                }
                uint32_t gtid = 0xffffffffu - (uint32_t)(key & 0xffffffffu);
                /* largest instance id with tri_prefix[id] <= gtid */
                uint32_t lo = 0, hi = n_inst;
                while (hi - lo > 1) {
                    uint32_t mid = (lo + hi) / 2;
                    if (tri_prefix[mid] <= gtid) lo = mid; else hi = mid;
                }
                uint32_t ii = lo;
                const tr_primitive_info* prim = &prims[inst[ii].primitive_id];
                tri_setup s;
                float l[3], d;
                int ok = setup_triangle(mesh, &inst[ii], prim, gtid - tri_prefix[ii], &pv, width, height, y0, y1, &s);
                ok = ok && eval_pixel(&s, (int)x, (int)yy, l, &d);
                (void)ok;
                v4 rot = v4_new(inst[ii].transform.rotation.x, inst[ii].transform.rotation.y,
                                inst[ii].transform.rotation.z, inst[ii].transform.rotation.w);
                v3 n[3];
                for (int k = 0; k < 3; k++) {
                    const float* mn = &mesh->normals[s.vid[k] * 3];
                    n[k] = quat_mul_v3(rot, v3_new(mn[0], mn[1], mn[2])); /* lib.rs:356 */
                }
                depth[i] = d;
                normal[i * 3 + 0] = (l[0] * n[0].x + l[1] * n[1].x) + l[2] * n[2].x;
                normal[i * 3 + 1] = (l[0] * n[0].y + l[1] * n[1].y) + l[2] * n[2].y;
                normal[i * 3 + 2] = (l[0] * n[0].z + l[1] * n[1].z) + l[2] * n[2].z;
                const float *u0 = &mesh->uvs[s.vid[0] * 2], *u1 = &mesh->uvs[s.vid[1] * 2], *u2 = &mesh->uvs[s.vid[2] * 2];
                uv[i * 2 + 0] = (l[0] * u0[0] + l[1] * u1[0]) + l[2] * u2[0];
                uv[i * 2 + 1] = (l[0] * u0[1] + l[1] * u1[1]) + l[2] * u2[1];
                mat[i] = inst[ii].material_id;
                if (layer == 1 && scale1) scale1[i] = inst[ii].transform.translation_and_scale.w;
                if (duv || ddepth) { /* forward differences to (x+1, y) and (x, y+1) on this triangle's plane */
                    for (int k = 0; k < 2; k++) {
                        float ln[3], dn;
                        float du = 0.0f, dv = 0.0f, dd = 0.0f;
                        if (eval_plane(&s, (int)x + (k == 0), (int)yy + (k == 1), ln, &dn)) {
                            float un = (ln[0] * u0[0] + ln[1] * u1[0]) + ln[2] * u2[0];
                            float vn = (ln[0] * u0[1] + ln[1] * u1[1]) + ln[2] * u2[1];
                            du = un - uv[i * 2 + 0];
                            dv = vn - uv[i * 2 + 1];
                            dd = dn - d;
                        }
                        if (duv) { duv[i * 4 + k * 2] = du; duv[i * 4 + k * 2 + 1] = dv; }
                        if (ddepth) ddepth[i * 2 + k] = dd;
                    }
                }
            }
        }
    }
    free(vis[0]);
    free(vis[1]);
    free(tri_prefix);
}

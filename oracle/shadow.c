/* oracle/shadow.c — CPU restatement of the ray-queried shadows (TEST INFRASTRUCTURE ONLY; see oracle.h).
 *
 * Reference behaviour:
 *   shader/src/lighting.rs:97-125   trace_shadow_ray: RayFlags::NONE, cull mask 0xff, t_min 0.001, every candidate
 *                                   confirmed (no alpha test), any committed hit -> factor 0, else 1
 *   shader/src/lighting.rs:22-32    sun ray: direction uniforms.sun_dir, t_max 10 000 (transmissive pass, factor as is)
 *   shader/src/lighting.rs:154-165  sun ray of the opaque pass, factor.max(0.1)
 *   shader/src/lighting.rs:64-71, 186-195   one ray per clustered light: direction and t_max = distance from
 *                                   light_direction_and_attenuation
 *   src/acceleration_structures.rs  one bottom-level structure per primitive (its index range, :47-52), one top-level
 *                                   structure over the instances
 *   src/main.rs:614-625             ... only the instances whose primitive has draw_buffer_index < 2 (opaque and
 *                                   alpha-clip geometry casts shadows, transmissive geometry does not)
 *   src/main.rs:2734-2754           instance transform = the instance's Similarity
 *
 * The ray/triangle arithmetic of VK_KHR_ray_query is implementation-defined (hardware in the reference): parity of WHICH
 * TRIANGLE A RAY MEETS is not pinnable, as for the rasteriser — that definition is ours (oracle.h).  What the shaders do around
 * the query (when a ray is traced, its parameters, what its answer does to the light: oracle/shade.c) is pinned to the
 * reference's ray-tracing builds of the fragment modules, compiled-shaders/ray-tracing/{fragment,fragment_transmission}.spv,
 * which run with orc_trace_shadow as their ray-query environment (oracle/spv_harness.c; tests/test_reference_spirv.py
 * test_live_ray_tracing_fragments, test_golden_ray_tracing_fragments: bit-equal pixels).
 * The definition used on both sides (DESIGN.md "Ray-queried shadows"):
 *
 *   occluded(ray) :=  exists an instance i of the top-level set with  slab(world ray, world_box_i)  and a triangle k of
 *                     its primitive with  slab(object ray_i, box_k)  and  hit64(object ray_i, triangle k)
 *
 *   - world_box_i: the eight corners of the primitive's object-space box through the instance's Similarity (fp32, the
 *     operation order of `Similarity * Vec3`), min/max;
 *   - object ray_i: o' = (conj(q) * (o - t)) / s,  d' = (conj(q) * d) / s  in fp32, so t keeps its world meaning;
 *   - slab(): the fp32 interval test below; it is MONOTONE in the box (a larger box never fails where a smaller one
 *     passes), so any bounding-volume hierarchy whose inner boxes are exact unions of its leaves can be used to skip
 *     work without changing the result — which is why the result does not depend on how the product builds its trees;
 *   - hit64(): Moeller-Trumbore in double on the fp32 inputs, both faces, t_min < t < t_max.
 *
 * The trees here are deliberately not the product's (median split on the widest centroid axis against its binned SAH),
 * and orc_trace_shadow_brute() walks no tree at all; tests/ compare all three.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "vecmath.h"

typedef struct {
    float lo[3], hi[3];
} box3;

typedef struct {
    box3 b;
    int left, right; /* inner node */
    int first, count; /* leaf when count > 0: items[first .. first+count) */
} bnode;

typedef struct {
    bnode* nodes;
    int n_nodes, cap;
    int* items;
} tree;

typedef struct {
    v3 translation;
    float inv_scale;
    v4 conj_rotation;
    uint32_t primitive;
    box3 world;
} accel_instance;

struct orc_accel {
    uint32_t n_prims;
    uint32_t* tri_first;  /* [n_prims+1] into tris */
    float* tris;          /* 9 floats per triangle */
    box3* tri_box;
    box3* prim_box;
    tree* blas;           /* [n_prims], items are triangle numbers within the primitive */
    uint32_t n_instances; /* top-level set */
    accel_instance* inst;
    tree tlas;
};

static box3 box_empty(void) {
    box3 b = {{INFINITY, INFINITY, INFINITY}, {-INFINITY, -INFINITY, -INFINITY}};
    return b;
}
static void box_grow(box3* b, const float p[3]) {
    for (int k = 0; k < 3; k++) {
        b->lo[k] = fminf(b->lo[k], p[k]);
        b->hi[k] = fmaxf(b->hi[k], p[k]);
    }
}
static void box_union(box3* b, const box3* o) {
    box_grow(b, o->lo);
    box_grow(b, o->hi);
}

/* ---- the two predicates of the definition ------------------------------------------------------------------------ */
typedef struct {
    float o[3], d[3], inv[3];
    float t_min, t_max;
} ray;

static ray make_ray(v3 o, v3 d, float t_min, float t_max) {
    ray r = {{o.x, o.y, o.z}, {d.x, d.y, d.z}, {1.0f / d.x, 1.0f / d.y, 1.0f / d.z}, t_min, t_max};
    return r;
}

static int slab(const ray* r, const box3* b) {
    float tn = r->t_min, tf = r->t_max;
    for (int k = 0; k < 3; k++) {
        int neg = r->inv[k] < 0.0f;
        float a = ((neg ? b->hi[k] : b->lo[k]) - r->o[k]) * r->inv[k];
        float c = ((neg ? b->lo[k] : b->hi[k]) - r->o[k]) * r->inv[k];
        tn = fmaxf(tn, a); /* NaN (0 * inf: origin on a face, direction parallel to it) does not constrain */
        tf = fminf(tf, c);
    }
    return tn <= tf * 1.00000024f; /* 1 + 2^-22 */
}

static int hit64(const ray* r, const float* v) {
    double e1[3], e2[3], tv[3], p[3], q[3];
    for (int k = 0; k < 3; k++) {
        e1[k] = (double)v[3 + k] - (double)v[k];
        e2[k] = (double)v[6 + k] - (double)v[k];
        tv[k] = (double)r->o[k] - (double)v[k];
    }
    const double d[3] = {r->d[0], r->d[1], r->d[2]};
    p[0] = d[1] * e2[2] - d[2] * e2[1];
    p[1] = d[2] * e2[0] - d[0] * e2[2];
    p[2] = d[0] * e2[1] - d[1] * e2[0];
    double det = (e1[0] * p[0] + e1[1] * p[1]) + e1[2] * p[2];
    if (det == 0.0) return 0;
    double inv = 1.0 / det;
    double u = ((tv[0] * p[0] + tv[1] * p[1]) + tv[2] * p[2]) * inv;
    if (!(u >= 0.0 && u <= 1.0)) return 0;
    q[0] = tv[1] * e1[2] - tv[2] * e1[1];
    q[1] = tv[2] * e1[0] - tv[0] * e1[2];
    q[2] = tv[0] * e1[1] - tv[1] * e1[0];
    double w = ((d[0] * q[0] + d[1] * q[1]) + d[2] * q[2]) * inv;
    if (!(w >= 0.0 && u + w <= 1.0)) return 0;
    double t = ((e2[0] * q[0] + e2[1] * q[1]) + e2[2] * q[2]) * inv;
    return t > (double)r->t_min && t < (double)r->t_max;
}

static ray object_ray(const accel_instance* in, v3 o, v3 d, float t_min, float t_max) {
    v3 oo = v3_scale(quat_mul_v3(in->conj_rotation, v3_sub(o, in->translation)), in->inv_scale);
    v3 od = v3_scale(quat_mul_v3(in->conj_rotation, d), in->inv_scale);
    return make_ray(oo, od, t_min, t_max);
}

/* ---- median-split trees ------------------------------------------------------------------------------------------ */
static const box3* g_sort_boxes;
static int g_sort_axis;
static int cmp_centroid(const void* a, const void* b) {
    const box3 *x = &g_sort_boxes[*(const int*)a], *y = &g_sort_boxes[*(const int*)b];
    float cx = x->lo[g_sort_axis] + x->hi[g_sort_axis], cy = y->lo[g_sort_axis] + y->hi[g_sort_axis];
    if (cx < cy) return -1;
    if (cx > cy) return 1;
    return *(const int*)a - *(const int*)b;
}

static int tree_node(tree* t) {
    if (t->n_nodes == t->cap) {
        t->cap = t->cap ? t->cap * 2 : 64;
        t->nodes = (bnode*)realloc(t->nodes, (size_t)t->cap * sizeof(bnode));
    }
    return t->n_nodes++;
}

static int tree_build_range(tree* t, const box3* boxes, int first, int count, int leaf_max) {
    int id = tree_node(t);
    box3 b = box_empty(), cb = box_empty();
    for (int i = 0; i < count; i++) {
        const box3* x = &boxes[t->items[first + i]];
        box_union(&b, x);
        float c[3] = {x->lo[0] + x->hi[0], x->lo[1] + x->hi[1], x->lo[2] + x->hi[2]};
        box_grow(&cb, c);
    }
    t->nodes[id].b = b;
    t->nodes[id].left = t->nodes[id].right = -1;
    t->nodes[id].first = first;
    t->nodes[id].count = count;
    if (count <= leaf_max) return id;
    int axis = 0;
    float ext = cb.hi[0] - cb.lo[0];
    for (int k = 1; k < 3; k++)
        if (cb.hi[k] - cb.lo[k] > ext) {
            ext = cb.hi[k] - cb.lo[k];
            axis = k;
        }
    g_sort_boxes = boxes;
    g_sort_axis = axis;
    qsort(t->items + first, (size_t)count, sizeof(int), cmp_centroid);
    int half = count / 2;
    int l = tree_build_range(t, boxes, first, half, leaf_max);
    int r = tree_build_range(t, boxes, first + half, count - half, leaf_max);
    t->nodes[id].left = l;
    t->nodes[id].right = r;
    t->nodes[id].count = 0;
    return id;
}

static void tree_build(tree* t, const box3* boxes, int n, int leaf_max) {
    memset(t, 0, sizeof(*t));
    t->items = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) t->items[i] = i;
    if (n > 0) tree_build_range(t, boxes, 0, n, leaf_max);
}

/* ---- build ------------------------------------------------------------------------------------------------------- */
orc_accel* orc_accel_build(const orc_mesh* mesh, const tr_instance* inst, uint32_t n_inst, const tr_primitive_info* prims,
                           uint32_t n_prims) {
    orc_accel* a = (orc_accel*)calloc(1, sizeof(orc_accel));
    a->n_prims = n_prims;
    a->tri_first = (uint32_t*)calloc(n_prims + 1, sizeof(uint32_t));
    for (uint32_t p = 0; p < n_prims; p++) a->tri_first[p + 1] = a->tri_first[p] + prims[p].index_count / 3;
    uint32_t n_tris = a->tri_first[n_prims];
    a->tris = (float*)malloc(sizeof(float) * 9 * (size_t)(n_tris ? n_tris : 1));
    a->tri_box = (box3*)malloc(sizeof(box3) * (size_t)(n_tris ? n_tris : 1));
    a->prim_box = (box3*)malloc(sizeof(box3) * (size_t)(n_prims ? n_prims : 1));
    a->blas = (tree*)calloc(n_prims ? n_prims : 1, sizeof(tree));
    for (uint32_t p = 0; p < n_prims; p++) { /* one bottom-level structure per primitive, acceleration_structures.rs:47-52 */
        a->prim_box[p] = box_empty();
        uint32_t nt = prims[p].index_count / 3;
        for (uint32_t t = 0; t < nt; t++) {
            uint32_t g = a->tri_first[p] + t;
            a->tri_box[g] = box_empty();
            for (int c = 0; c < 3; c++) {
                uint32_t vi = mesh->indices[prims[p].first_index + t * 3 + c];
                memcpy(&a->tris[(size_t)g * 9 + c * 3], &mesh->positions[(size_t)vi * 3], 12);
                box_grow(&a->tri_box[g], &a->tris[(size_t)g * 9 + c * 3]);
            }
            box_union(&a->prim_box[p], &a->tri_box[g]);
        }
        tree_build(&a->blas[p], a->tri_box + a->tri_first[p], (int)nt, 2);
    }
    a->inst = (accel_instance*)malloc(sizeof(accel_instance) * (size_t)(n_inst ? n_inst : 1));
    box3* wb = (box3*)malloc(sizeof(box3) * (size_t)(n_inst ? n_inst : 1));
    for (uint32_t i = 0; i < n_inst; i++) {
        const tr_primitive_info* pr = &prims[inst[i].primitive_id];
        if (pr->draw_buffer_index >= 2u || pr->index_count < 3u) continue; /* src/main.rs:617-620 */
        accel_instance* in = &a->inst[a->n_instances];
        const tr_packed_similarity* s = &inst[i].transform;
        v3 tr = v3_new(s->translation_and_scale.x, s->translation_and_scale.y, s->translation_and_scale.z);
        v4 q = v4_new(s->rotation.x, s->rotation.y, s->rotation.z, s->rotation.w);
        float scale = s->translation_and_scale.w;
        in->translation = tr;
        in->inv_scale = 1.0f / scale;
        in->conj_rotation = v4_new(-q.x, -q.y, -q.z, q.w);
        in->primitive = inst[i].primitive_id;
        in->world = box_empty();
        const box3* ob = &a->prim_box[in->primitive];
        for (int c = 0; c < 8; c++) {
            v3 corner = v3_new(c & 1 ? ob->hi[0] : ob->lo[0], c & 2 ? ob->hi[1] : ob->lo[1], c & 4 ? ob->hi[2] : ob->lo[2]);
            v3 w = v3_add(tr, v3_scale(quat_mul_v3(q, corner), scale)); /* Similarity * Vec3, shared-structs lib.rs:216-220 */
            float wp[3] = {w.x, w.y, w.z};
            box_grow(&in->world, wp);
        }
        wb[a->n_instances++] = in->world;
    }
    tree_build(&a->tlas, wb, (int)a->n_instances, 1);
    free(wb);
    return a;
}

void orc_accel_free(orc_accel* a) {
    if (!a) return;
    for (uint32_t p = 0; p < a->n_prims; p++) {
        free(a->blas[p].nodes);
        free(a->blas[p].items);
    }
    free(a->tlas.nodes);
    free(a->tlas.items);
    free(a->blas);
    free(a->tri_first);
    free(a->tris);
    free(a->tri_box);
    free(a->prim_box);
    free(a->inst);
    free(a);
}

uint32_t orc_accel_instance_count(const orc_accel* a) { return a->n_instances; }

/* ---- trace ------------------------------------------------------------------------------------------------------- */
static int instance_occludes(const orc_accel* a, const accel_instance* in, v3 o, v3 d, float t_min, float t_max) {
    ray r = object_ray(in, o, d, t_min, t_max);
    const tree* t = &a->blas[in->primitive];
    if (t->n_nodes == 0) return 0;
    uint32_t base = a->tri_first[in->primitive];
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const bnode* n = &t->nodes[stack[--sp]];
        if (!slab(&r, &n->b)) continue;
        if (n->count > 0) {
            for (int i = 0; i < n->count; i++) {
                uint32_t g = base + (uint32_t)t->items[n->first + i];
                if (slab(&r, &a->tri_box[g]) && hit64(&r, &a->tris[(size_t)g * 9])) return 1;
            }
        } else {
            stack[sp++] = n->left;
            stack[sp++] = n->right;
        }
    }
    return 0;
}

/* trace_shadow_ray, lighting.rs:97-125: 1.0 = lit, 0.0 = occluded */
float orc_trace_shadow(const orc_accel* a, v3 origin, v3 direction, float t_max) {
    const float t_min = 0.001f;
    if (a->tlas.n_nodes == 0) return 1.0f;
    ray wr = make_ray(origin, direction, t_min, t_max);
    int stack[128], sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const bnode* n = &a->tlas.nodes[stack[--sp]];
        if (!slab(&wr, &n->b)) continue;
        if (n->count > 0) {
            for (int i = 0; i < n->count; i++) {
                const accel_instance* in = &a->inst[a->tlas.items[n->first + i]];
                if (slab(&wr, &in->world) && instance_occludes(a, in, origin, direction, t_min, t_max)) return 0.0f;
            }
        } else {
            stack[sp++] = n->left;
            stack[sp++] = n->right;
        }
    }
    return 1.0f;
}

/* the definition without any tree: every instance, every triangle */
float orc_trace_shadow_brute(const orc_accel* a, v3 origin, v3 direction, float t_max) {
    const float t_min = 0.001f;
    ray wr = make_ray(origin, direction, t_min, t_max);
    for (uint32_t i = 0; i < a->n_instances; i++) {
        const accel_instance* in = &a->inst[i];
        if (!slab(&wr, &in->world)) continue;
        ray r = object_ray(in, origin, direction, t_min, t_max);
        uint32_t g0 = a->tri_first[in->primitive], g1 = a->tri_first[in->primitive + 1];
        for (uint32_t g = g0; g < g1; g++)
            if (slab(&r, &a->tri_box[g]) && hit64(&r, &a->tris[(size_t)g * 9])) return 0.0f;
    }
    return 1.0f;
}

void orc_trace_shadow_rays(const orc_accel* a, uint32_t n, const float* origins, const float* directions, const float* t_max,
                           int brute, uint8_t* lit) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        v3 o = v3_new(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]);
        v3 d = v3_new(directions[i * 3], directions[i * 3 + 1], directions[i * 3 + 2]);
        float f = brute ? orc_trace_shadow_brute(a, o, d, t_max[i]) : orc_trace_shadow(a, o, d, t_max[i]);
        lit[i] = f != 0.0f;
    }
}

/* the slab predicate on its own, for the property tests (monotone in the box: tests/test_oracle_shadows.py) */
int orc_slab_test(const float origin[3], const float direction[3], float t_min, float t_max, const float lo[3], const float hi[3]) {
    ray r = make_ray(v3_new(origin[0], origin[1], origin[2]), v3_new(direction[0], direction[1], direction[2]), t_min, t_max);
    box3 b = {{lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}};
    return slab(&r, &b);
}

// k_clusters.cu — K2a cluster AABBs and K2b light -> cluster assignment (sm_100a).
//
// Reference: `write_cluster_data` shader/src/lib.rs:519-594 (4x4x4 groups, main.rs:1511-1516) and
// `assign_lights_to_clusters` lib.rs:596-645 (one thread per (cluster, light) pair appending
// through an atomic counter, main.rs:1792-1797), with ClusterAabb::{distance_sq, cull_spotlight}
// shared-structs/src/lib.rs:290-320 and slice_to_depth :65-67.
// Membership is a discrete decision => exact regime throughout; the transcendental inputs
// (powf in slice_to_depth, cos/sin in cull_spotlight) are evaluated in double and rounded once,
// like the oracle.  K2b gives one WARP to each cluster: lanes test 32 lights at a time and
// append with ballot + popc, so every list comes out in ascending light id (the reference's
// atomic order is a race) and no atomics are needed at all.
#include "tr_internal.h"

using namespace trd;

namespace {

struct ClusterParams {
    tr_uniforms u;
    tr_write_cluster_data_push_constants pc;
    uint32_t nz;
    tr_cluster_aabb* out;
};

// shared-structs lib.rs:65-67
__device__ __forceinline__ float slice_to_depth(const tr_light_cluster_coefficients& c, uint32_t slice) {
    const float ratio = xdiv(c.z_far, c.z_near);
    const float t = xdiv((float)slice, (float)c.num_depth_slices);
    return xmul(-c.z_near, (float)pow((double)ratio, (double)t));
}

// shader/src/lib.rs:583-594
__device__ __forceinline__ f3 line_intersection_to_z_plane(f3 a, f3 b, float z_distance) {
    const f3 normal = mk3(0.0f, 0.0f, 1.0f);
    const f3 a_to_b = xsub3(b, a);
    const float t = xdiv(xsub(z_distance, xdot3(normal, a)), xdot3(normal, a_to_b));
    return xadd3(a, xscale3(a_to_b, t));
}

__device__ __forceinline__ f3 min3(f3 a, f3 b) { return mk3(rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z)); }
__device__ __forceinline__ f3 max3(f3 a, f3 b) { return mk3(rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z)); }

__global__ void __launch_bounds__(128) cluster_aabb_kernel(const __grid_constant__ ClusterParams p) {
    const uint32_t nx = p.u.num_clusters.x, ny = p.u.num_clusters.y;
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nx * ny * p.nz) return;
    const uint32_t iz = id / (nx * ny), rem = id - iz * nx * ny, iy = rem / nx, ix = rem - iy * nx;
    const mat4& inv_persp = *reinterpret_cast<const mat4*>(&p.pc.inverse_perspective);
    f3 view_space[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        // screen_space_{min,max} :537-538, screen_to_clip :540-544, clip_to_view :546-550
        float sx = xmul((float)(ix + k), p.u.cluster_size_in_pixels.x);
        float sy = xmul((float)(iy + k), p.u.cluster_size_in_pixels.y);
        float cx = xsub(xmul(xdiv(sx, (float)p.pc.screen_dimensions.x), 2.0f), 1.0f);
        float cy = xsub(xmul(xdiv(sy, (float)p.pc.screen_dimensions.y), 2.0f), 1.0f);
        f4 v = xmat4_mul(inv_persp, cx, cy, 0.0f, 1.0f);
        view_space[k] = xdivs3(mk3(v.x, v.y, v.z), v.w);
    }
    const float z_near = slice_to_depth(p.u.light_clustering_coefficients, iz);
    const float z_far = slice_to_depth(p.u.light_clustering_coefficients, iz + 1);
    const f3 eye = mk3(0.0f, 0.0f, 1.0f);  // :560 (sic)
    const f3 a = line_intersection_to_z_plane(eye, view_space[0], z_near);
    const f3 b = line_intersection_to_z_plane(eye, view_space[0], z_far);
    const f3 c = line_intersection_to_z_plane(eye, view_space[1], z_near);
    const f3 d = line_intersection_to_z_plane(eye, view_space[1], z_far);
    const f3 lo = min3(min3(min3(a, b), c), d), hi = max3(max3(max3(a, b), c), d);
    tr_cluster_aabb o;
    o.min.x = lo.x; o.min.y = lo.y; o.min.z = lo.z; o.min._pad = 0.0f;
    o.max.x = hi.x; o.max.y = hi.y; o.max.z = hi.z; o.max._pad = 0.0f;
    p.out[id] = o;
}

struct AssignParams {
    const tr_light* lights;
    uint32_t n_lights;
    const tr_cluster_aabb* clusters;
    uint32_t n_clusters;
    tr_assign_lights_push_constants pc;
    uint32_t* counts;
    uint32_t* indices;
};

// shared-structs lib.rs:291-298
__device__ __forceinline__ float aabb_distance_sq(f3 lo, f3 hi, f3 p) {
    f3 d = max3(max3(xsub3(lo, p), xsub3(p, hi)), mk3(0.0f, 0.0f, 0.0f));
    return xdot3(d, d);
}

// shared-structs lib.rs:301-319
__device__ __forceinline__ bool aabb_cull_spotlight(f3 lo, f3 hi, f3 origin, f3 direction, float angle, float range) {
    const f3 center = xdivs3(xadd3(lo, hi), 2.0f);
    const f3 dc = xsub3(hi, center);
    const float radius = xsqrt(xdot3(dc, dc));
    const f3 vector = xsub3(center, origin);
    const float vector_len_sq = xdot3(vector, vector);
    const float v1 = xdot3(vector, direction);
    const float v1sq = xmul(v1, v1);
    const float ca = (float)cos((double)angle), sa = (float)sin((double)angle);
    const float closest = xsub(xmul(ca, xsqrt(xsub(vector_len_sq, v1sq))), xmul(v1, sa));
    const bool angle_cull = closest > radius;
    const bool front_cull = v1 > xadd(radius, range);
    const bool back_cull = v1 < -radius;
    return angle_cull || front_cull || back_cull;
}

__global__ void __launch_bounds__(256) assign_lights_kernel(const __grid_constant__ AssignParams p) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t cluster = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cluster >= p.n_clusters) return;
    const float4* cq = reinterpret_cast<const float4*>(p.clusters + cluster);
    const float4 cmin = __ldg(cq), cmax = __ldg(cq + 1);
    const f3 lo = mk3(cmin.x, cmin.y, cmin.z), hi = mk3(cmax.x, cmax.y, cmax.z);
    const mat4& view = *reinterpret_cast<const mat4*>(&p.pc.view_matrix);
    uint32_t n = 0;
    for (uint32_t base = 0; base < p.n_lights; base += 32) {
        const uint32_t l = base + lane;
        bool keep = false;
        if (l < p.n_lights) {
            const float4* lq = reinterpret_cast<const float4*>(p.lights + l);
            const float4 pos = __ldg(lq), col = __ldg(lq + 1), spot = __ldg(lq + 2);
            const f4 lp4 = xmat4_mul(view, pos.x, pos.y, pos.z, 1.0f);  // lib.rs:620
            const f3 lp = mk3(lp4.x, lp4.y, lp4.z);
            keep = !(aabb_distance_sq(lo, hi, lp) > col.w);             // lib.rs:622-626
            if (keep && spot.w != 0.0f) {                               // lib.rs:628-638
                const f3 dir = xquat_mul3(p.pc.view_rotation.x, p.pc.view_rotation.y, p.pc.view_rotation.z,
                                          p.pc.view_rotation.w, mk3(spot.x, spot.y, spot.z));
                if (aabb_cull_spotlight(lo, hi, lp, dir, spot.w, col.w)) keep = false;
            }
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const uint32_t slot = n + __popc(mask & ((1u << lane) - 1u));
            // the reference has no bound check (a 129th light would spill into the next cluster's
            // slots, lib.rs:640-644); we drop it and saturate the count, like the oracle.
            if (slot < TR_MAX_LIGHTS_PER_CLUSTER) p.indices[(size_t)cluster * TR_MAX_LIGHTS_PER_CLUSTER + slot] = l;
        }
        n += __popc(mask);
    }
    if (lane == 0) p.counts[cluster] = n < TR_MAX_LIGHTS_PER_CLUSTER ? n : TR_MAX_LIGHTS_PER_CLUSTER;
}

}  // namespace

namespace tr {

int32_t launch_build_clusters(tr_ctx* c, const tr_write_cluster_data_push_constants& pc) {
    if (!c->have_uniforms) return fail(TR_ERR_STATE, "tr_build_clusters: uniforms not set");
    ClusterParams p;
    p.u = c->uniforms;
    p.pc = pc;
    p.nz = c->uniforms.light_clustering_coefficients.num_depth_slices;
    TR_TRY(c->cluster_aabbs.ensure((size_t)c->n_clusters * sizeof(tr_cluster_aabb)));
    p.out = c->cluster_aabbs.as<tr_cluster_aabb>();
    cluster_aabb_kernel<<<(c->n_clusters + 127) / 128, 128, 0, c->stream>>>(p);
    count_launches(1);
    TR_CUDA(cudaGetLastError());
    c->clusters_valid = true;
    return TR_OK;
}

int32_t launch_assign_lights(tr_ctx* c, const tr_assign_lights_push_constants& pc) {
    if (!c->clusters_valid) return fail(TR_ERR_STATE, "tr_assign_lights: cluster AABBs not built (tr_build_clusters)");
    TR_TRY(c->cluster_counts.ensure((size_t)c->n_clusters * 4));
    TR_TRY(c->cluster_indices.ensure((size_t)c->n_clusters * TR_MAX_LIGHTS_PER_CLUSTER * 4));
    AssignParams p;
    p.lights = c->lights.as<tr_light>();
    p.n_lights = c->n_lights;
    p.clusters = c->cluster_aabbs.as<tr_cluster_aabb>();
    p.n_clusters = c->n_clusters;
    p.pc = pc;
    p.counts = c->cluster_counts.as<uint32_t>();
    p.indices = c->cluster_indices.as<uint32_t>();
    const uint32_t threads = 256, warps_per_block = threads / 32;
    assign_lights_kernel<<<(c->n_clusters + warps_per_block - 1) / warps_per_block, threads, 0, c->stream>>>(p);
    count_launches(1);
    TR_CUDA(cudaGetLastError());
    c->cluster_lights_valid = true;
    return TR_OK;
}

}  // namespace tr

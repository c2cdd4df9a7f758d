// k_visibility.cu — K3: deterministic depth pre-pass / visibility + G-buffer resolve (sm_100a).
//
// Stands in for the reference's hardware rasteriser and varying interpolation:
//   depth_pre_pass_instanced  shader/src/lib.rs:319-333   (clip = proj_view * (T * pos))
//   vertex_instanced[_with_scale] lib.rs:335-391          (rotation * normal, uv, material_id, scale)
//   pipeline state src/pipelines.rs:311,350-371 (back-face cull, depth GREATER + write, reversed-Z),
//   clear 0.0 src/main.rs:1586-1591, pass order src/main.rs:1900-1944 / 2005-2042.
// The rules are ours (DESIGN.md "visibility", oracle/raster.c is the same algorithm on the CPU and
// the two agree bit for bit): homogeneous edge functions in double so triangles crossing the camera
// plane need no clipping, pixel centres at +0.5, CCW-from-outside front faces, an exact shared-edge
// tie rule, nearest fragment wins through a 64-bit atomicMax on (depth bits << 32 | ~triangle id)
// — order independent, hence deterministic —, transmissive layer tested GREATER against the final
// opaque depth.  Work list = the visible instances' triangles from K1's scan (no host round trip).
//   pass 1  one thread per triangle; small bounding boxes are rasterised in place, large ones are
//           cut into 64x64-pixel tiles (conservatively culled against the edges) and queued
//   pass 2  one warp per queued tile
//   pass 3  per pixel: re-evaluate the winning triangle, interpolate perspective-correct varyings,
//           write the SoA G-buffer planes of both layers, and clear the visibility words for the
//           next frame.
#include "tr_internal.h"

using namespace trd;

namespace {

constexpr int RASTER_THREADS = 128;
constexpr int SMALL_BBOX_PIXELS = 256;
constexpr int BIG_TILE = 64;
constexpr uint32_t QUEUE_CAPACITY = 1u << 20;

struct VisParams {
    const float* positions;
    const float* normals;
    const float* uvs;
    const uint32_t* indices;
    const tr_instance* instances;
    const tr_primitive_info* prims;
    const uint32_t* visible_ids;
    const uint32_t* work_prefix;  // [n_visible + 1]
    const uint32_t* scalars;      // [0] n_visible, [1] total triangles
    mat4 proj_view;
    uint32_t width, height, y0, y1;
    unsigned long long* vis[2];
    uint4* queue;                 // (slot, tri, tile_x | tile_y << 16, layer)
    uint32_t* queue_count;        // [2]: per layer
    // resolve outputs
    float* depth[2];
    float* normal[2];
    float* uv[2];
    uint32_t* material_id[2];
    float* scale1;
};

struct TriSetup {
    double A[3], B[3], C[3];
    float Z[3], W[3];
    uint32_t vid[3];
    int x_lo, x_hi, y_lo, y_hi;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// oracle/raster.c setup_triangle
__device__ __forceinline__ bool setup_triangle(const VisParams& p, const tr_instance* inst, const tr_primitive_info* prim,
                                               uint32_t tri, TriSetup& s) {
    const float4* iq = reinterpret_cast<const float4*>(inst);
    const float4 ts = __ldg(iq), rot = __ldg(iq + 1);
    const uint32_t first_index = __ldg(&prim->first_index);
    uint32_t vid[3];
    float sx[3], sy[3], Z[3], W[3];
    const float half_w = xmul((float)p.width, 0.5f), half_h = xmul((float)p.height, 0.5f);
    bool finite = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        vid[k] = __ldg(p.indices + first_index + tri * 3 + k);
        const f3 pos = mk3(__ldg(p.positions + vid[k] * 3), __ldg(p.positions + vid[k] * 3 + 1), __ldg(p.positions + vid[k] * 3 + 2));
        const f3 wp = xadd3(mk3(ts.x, ts.y, ts.z), xscale3(xquat_mul3(rot.x, rot.y, rot.z, rot.w, pos), ts.w));
        const f4 c = xmat4_mul(p.proj_view, wp.x, wp.y, wp.z, 1.0f);
        finite = finite && isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && isfinite(c.w);
        sx[k] = xmul(xadd(c.x, c.w), half_w);
        sy[k] = xmul(xadd(c.y, c.w), half_h);
        Z[k] = c.z;
        W[k] = c.w;
    }
    if (!finite) return false;
    const double a0 = dsub(dmul(sy[1], W[2]), dmul(W[1], sy[2]));
    const double b0 = dsub(dmul(W[1], sx[2]), dmul(sx[1], W[2]));
    const double c0 = dsub(dmul(sx[1], sy[2]), dmul(sy[1], sx[2]));
    const double det = dadd(dadd(dmul(sx[0], a0), dmul(sy[0], b0)), dmul(W[0], c0));
    if (!(det < 0.0)) return false;  // back face / degenerate
    const int order[3] = {0, 2, 1};
    float rx[3], ry[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        rx[k] = sx[order[k]];
        ry[k] = sy[order[k]];
        s.Z[k] = Z[order[k]];
        s.W[k] = W[order[k]];
        s.vid[k] = vid[order[k]];
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        s.A[i] = dsub(dmul(ry[a], s.W[b]), dmul(s.W[a], ry[b]));
        s.B[i] = dsub(dmul(s.W[a], rx[b]), dmul(rx[a], s.W[b]));
        s.C[i] = dsub(dmul(rx[a], ry[b]), dmul(ry[a], rx[b]));
    }
    int x_lo = 0, x_hi = (int)p.width - 1, y_lo = (int)p.y0, y_hi = (int)p.y1 - 1;
    if (s.W[0] > 0.0f && s.W[1] > 0.0f && s.W[2] > 0.0f) {
        float px[3], py[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            px[k] = xdiv(rx[k], s.W[k]);
            py[k] = xdiv(ry[k], s.W[k]);
        }
        const float mnx = rmin(px[0], rmin(px[1], px[2])), mxx = rmax(px[0], rmax(px[1], px[2]));
        const float mny = rmin(py[0], rmin(py[1], py[2])), mxy = rmax(py[0], rmax(py[1], py[2]));
        if (!(mxx >= 0.0f) || !(mnx <= (float)p.width) || !(mxy >= (float)p.y0) || !(mny <= (float)p.y1)) return false;
        const float fx_lo = floorf(xsub(mnx, 0.5f)), fx_hi = ceilf(xsub(mxx, 0.5f));
        const float fy_lo = floorf(xsub(mny, 0.5f)), fy_hi = ceilf(xsub(mxy, 0.5f));
        if (fx_lo > (float)x_lo) x_lo = (int)fx_lo;
        if (fx_hi < (float)x_hi) x_hi = (int)fx_hi;
        if (fy_lo > (float)y_lo) y_lo = (int)fy_lo;
        if (fy_hi < (float)y_hi) y_hi = (int)fy_hi;
    }
    if (x_lo > x_hi || y_lo > y_hi) return false;
    s.x_lo = x_lo; s.x_hi = x_hi; s.y_lo = y_lo; s.y_hi = y_hi;
    return true;
}

// oracle/raster.c eval_pixel
__device__ __forceinline__ bool eval_pixel(const TriSetup& s, int px, int py, float l[3], float& depth) {
    const double qx = (double)px + 0.5, qy = (double)py + 0.5;
    double E[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        E[i] = dadd(dadd(dmul(s.A[i], qx), dmul(s.B[i], qy)), s.C[i]);
        if (E[i] < 0.0) return false;
        if (E[i] == 0.0 && !(s.A[i] > 0.0 || (s.A[i] == 0.0 && s.B[i] > 0.0))) return false;
    }
    const double S = dadd(dadd(E[0], E[1]), E[2]);
    if (!(S > 0.0)) return false;
    l[0] = __double2float_rn(__ddiv_rn(E[0], S));
    l[1] = __double2float_rn(__ddiv_rn(E[1], S));
    l[2] = __double2float_rn(__ddiv_rn(E[2], S));
    const float zq = xadd(xadd(xmul(l[0], s.Z[0]), xmul(l[1], s.Z[1])), xmul(l[2], s.Z[2]));
    const float wq = xadd(xadd(xmul(l[0], s.W[0]), xmul(l[1], s.W[1])), xmul(l[2], s.W[2]));
    const float d = xdiv(zq, wq);
    if (!(d > 0.0f) || d > 1.0f) return false;
    depth = d;
    return true;
}

__device__ __forceinline__ void plot(const VisParams& p, int layer, int px, int py, float d, uint32_t gtid) {
    const size_t i = (size_t)py * p.width + px;
    if (layer == 1) {  // depth GREATER against the opaque depth already in the shared depth buffer
        const float dop = __uint_as_float((uint32_t)(p.vis[0][i] >> 32));
        if (!(d > dop)) return;
    }
    const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(0xffffffffu - gtid);
    if (p.vis[layer][i] < key) atomicMax(p.vis[layer] + i, key);
}

// conservative: can the tile [x0,x1] x [y0,y1] (pixel indices) contain a covered pixel centre?
__device__ __forceinline__ bool tile_may_overlap(const TriSetup& s, int x0, int y0, int x1, int y1) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double qx = (s.A[i] >= 0.0 ? (double)x1 : (double)x0) + 0.5;
        const double qy = (s.B[i] >= 0.0 ? (double)y1 : (double)y0) + 0.5;
        // evaluate with a safety margin of one pixel of slope so rounding can never reject a covered tile
        const double e = s.A[i] * qx + s.B[i] * qy + s.C[i];
        const double margin = fabs(s.A[i]) + fabs(s.B[i]);
        if (e + margin < 0.0) return false;
    }
    return true;
}

template <int LAYER>
__global__ void __launch_bounds__(RASTER_THREADS) raster_kernel(const __grid_constant__ VisParams p) {
    const uint32_t n_visible = p.scalars[0], total = p.scalars[1];
    const uint32_t bucket = LAYER == 0 ? 0u : 2u;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
        // slot = largest s with work_prefix[s] <= w
        uint32_t lo = 0, hi = n_visible;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(p.work_prefix + mid) <= w) lo = mid; else hi = mid;
        }
        const uint32_t slot = lo, tri = w - __ldg(p.work_prefix + slot);
        const tr_instance* inst = p.instances + __ldg(p.visible_ids + slot);
        const tr_primitive_info* prim = p.prims + __ldg(&inst->primitive_id);
        if (__ldg(&prim->draw_buffer_index) != bucket) continue;
        TriSetup s;
        if (!setup_triangle(p, inst, prim, tri, s)) continue;
        const int bw = s.x_hi - s.x_lo + 1, bh = s.y_hi - s.y_lo + 1;
        if ((long long)bw * bh <= SMALL_BBOX_PIXELS) {
            for (int py = s.y_lo; py <= s.y_hi; py++)
                for (int px = s.x_lo; px <= s.x_hi; px++) {
                    float l[3], d;
                    if (eval_pixel(s, px, py, l, d)) plot(p, LAYER, px, py, d, w);
                }
        } else {
            const int tx0 = s.x_lo / BIG_TILE, tx1 = s.x_hi / BIG_TILE, ty0 = s.y_lo / BIG_TILE, ty1 = s.y_hi / BIG_TILE;
            for (int ty = ty0; ty <= ty1; ty++)
                for (int tx = tx0; tx <= tx1; tx++) {
                    const int x0 = max(tx * BIG_TILE, s.x_lo), x1 = min(tx * BIG_TILE + BIG_TILE - 1, s.x_hi);
                    const int y0 = max(ty * BIG_TILE, s.y_lo), y1 = min(ty * BIG_TILE + BIG_TILE - 1, s.y_hi);
                    if (!tile_may_overlap(s, x0, y0, x1, y1)) continue;
                    const uint32_t q = atomicAdd(p.queue_count + LAYER, 1u);
                    if (q < QUEUE_CAPACITY) {
                        p.queue[(size_t)LAYER * QUEUE_CAPACITY + q] = make_uint4(slot, tri, (uint32_t)tx | ((uint32_t)ty << 16), w);
                    } else {  // queue full: rasterise the tile here (slow but always correct)
                        for (int py = y0; py <= y1; py++)
                            for (int px = x0; px <= x1; px++) {
                                float l[3], d;
                                if (eval_pixel(s, px, py, l, d)) plot(p, LAYER, px, py, d, w);
                            }
                    }
                }
        }
    }
}

template <int LAYER>
__global__ void __launch_bounds__(256) raster_tiles_kernel(const __grid_constant__ VisParams p) {
    const uint32_t n_items = min(p.queue_count[LAYER], QUEUE_CAPACITY);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += warps) {
        const uint4 q = p.queue[(size_t)LAYER * QUEUE_CAPACITY + item];
        const tr_instance* inst = p.instances + __ldg(p.visible_ids + q.x);
        const tr_primitive_info* prim = p.prims + __ldg(&inst->primitive_id);
        TriSetup s;
        if (!setup_triangle(p, inst, prim, q.y, s)) continue;
        const int tx = q.z & 0xffff, ty = q.z >> 16;
        const int x0 = max(tx * BIG_TILE, s.x_lo), x1 = min(tx * BIG_TILE + BIG_TILE - 1, s.x_hi);
        const int y0 = max(ty * BIG_TILE, s.y_lo), y1 = min(ty * BIG_TILE + BIG_TILE - 1, s.y_hi);
        const int tw = x1 - x0 + 1, n = tw * (y1 - y0 + 1);
        for (int i = lane; i < n; i += 32) {
            const int py = y0 + i / tw, px = x0 + i % tw;
            float l[3], d;
            if (eval_pixel(s, px, py, l, d)) plot(p, LAYER, px, py, d, q.w);
        }
    }
}

__global__ void __launch_bounds__(256) resolve_kernel(const __grid_constant__ VisParams p) {
    const uint32_t n_visible = p.scalars[0];
    const uint32_t begin = p.y0 * p.width, end = p.y1 * p.width;
    for (uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
        const uint32_t py = i / p.width, px = i - py * p.width;
#pragma unroll
        for (int layer = 0; layer < 2; layer++) {
            const unsigned long long key = p.vis[layer][i];
            if (key == 0ull) {
                p.depth[layer][i] = 0.0f;
                p.normal[layer][(size_t)i * 3] = 0.0f; p.normal[layer][(size_t)i * 3 + 1] = 0.0f; p.normal[layer][(size_t)i * 3 + 2] = 0.0f;
                p.uv[layer][(size_t)i * 2] = 0.0f; p.uv[layer][(size_t)i * 2 + 1] = 0.0f;
                p.material_id[layer][i] = 0xffffffffu;
                if (layer == 1) p.scale1[i] = 0.0f;
                continue;
            }
            p.vis[layer][i] = 0ull;  // clear for the next frame
            const uint32_t w = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
            uint32_t lo = 0, hi = n_visible;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(p.work_prefix + mid) <= w) lo = mid; else hi = mid;
            }
            const uint32_t tri = w - __ldg(p.work_prefix + lo);
            const tr_instance* inst = p.instances + __ldg(p.visible_ids + lo);
            const tr_primitive_info* prim = p.prims + __ldg(&inst->primitive_id);
            TriSetup s;
            float l[3] = {0.f, 0.f, 0.f}, d = 0.0f;
            const bool ok = setup_triangle(p, inst, prim, tri, s) && eval_pixel(s, (int)px, (int)py, l, d);
            (void)ok;  // by construction the winning triangle covers this pixel
            const float4 rot = __ldg(reinterpret_cast<const float4*>(inst) + 1);
            f3 n[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float* mn = p.normals + (size_t)s.vid[k] * 3;
                n[k] = xquat_mul3(rot.x, rot.y, rot.z, rot.w, mk3(__ldg(mn), __ldg(mn + 1), __ldg(mn + 2)));  // lib.rs:356
            }
            p.depth[layer][i] = d;
            p.normal[layer][(size_t)i * 3 + 0] = xadd(xadd(xmul(l[0], n[0].x), xmul(l[1], n[1].x)), xmul(l[2], n[2].x));
            p.normal[layer][(size_t)i * 3 + 1] = xadd(xadd(xmul(l[0], n[0].y), xmul(l[1], n[1].y)), xmul(l[2], n[2].y));
            p.normal[layer][(size_t)i * 3 + 2] = xadd(xadd(xmul(l[0], n[0].z), xmul(l[1], n[1].z)), xmul(l[2], n[2].z));
            const float *u0 = p.uvs + (size_t)s.vid[0] * 2, *u1 = p.uvs + (size_t)s.vid[1] * 2, *u2 = p.uvs + (size_t)s.vid[2] * 2;
            p.uv[layer][(size_t)i * 2 + 0] = xadd(xadd(xmul(l[0], __ldg(u0)), xmul(l[1], __ldg(u1))), xmul(l[2], __ldg(u2)));
            p.uv[layer][(size_t)i * 2 + 1] = xadd(xadd(xmul(l[0], __ldg(u0 + 1)), xmul(l[1], __ldg(u1 + 1))), xmul(l[2], __ldg(u2 + 1)));
            p.material_id[layer][i] = __ldg(&inst->material_id);
            if (layer == 1) p.scale1[i] = __ldg(&inst->transform.translation_and_scale.w);
        }
    }
}

}  // namespace

namespace tr {

int32_t launch_visibility(tr_ctx* c, const tr_push_constants& pc) {
    if (!c->cull_valid) return fail(TR_ERR_STATE, "tr_visibility: tr_cull has not run");
    if (!c->n_indices) return fail(TR_ERR_STATE, "tr_visibility: no mesh (tr_set_mesh)");
    const size_t npx = (size_t)c->width * c->height;
    for (int l = 0; l < 2; l++) {
        const bool fresh = c->vis[l].bytes < npx * 8;
        TR_TRY(c->vis[l].ensure(npx * 8));
        if (fresh) TR_CUDA(cudaMemsetAsync(c->vis[l].p, 0, npx * 8, c->stream));
        TR_TRY(ensure_layer(c, l, false));
    }
    const size_t queue_bytes = (size_t)2 * QUEUE_CAPACITY * sizeof(uint4);
    TR_TRY(c->big_queue.ensure(queue_bytes + 16));
    uint32_t* qcount = reinterpret_cast<uint32_t*>(c->big_queue.as<unsigned char>() + queue_bytes);
    TR_CUDA(cudaMemsetAsync(qcount, 0, 8, c->stream));

    VisParams p{};
    p.positions = c->mesh_pos.as<float>();
    p.normals = c->mesh_nrm.as<float>();
    p.uvs = c->mesh_uv.as<float>();
    p.indices = c->mesh_idx.as<uint32_t>();
    p.instances = c->instances.as<tr_instance>();
    p.prims = c->primitives.as<tr_primitive_info>();
    p.visible_ids = c->visible_ids.as<uint32_t>();
    p.work_prefix = c->work_prefix.as<uint32_t>();
    p.scalars = c->d_cull_scalars;
    memcpy(&p.proj_view, &pc.proj_view, sizeof(mat4));
    p.width = c->width;
    p.height = c->height;
    p.y0 = c->band_y0;
    p.y1 = c->band_y1;
    p.queue = c->big_queue.as<uint4>();
    p.queue_count = qcount;
    for (int l = 0; l < 2; l++) {
        p.vis[l] = c->vis[l].as<unsigned long long>();
        p.depth[l] = c->layer[l].depth.as<float>();
        p.normal[l] = c->layer[l].normal.as<float>();
        p.uv[l] = c->layer[l].uv.as<float>();
        p.material_id[l] = c->layer[l].material_id.as<uint32_t>();
    }
    p.scale1 = c->layer[1].scale.as<float>();

    const int grid = c->sm_count * 8;
    raster_kernel<0><<<grid, RASTER_THREADS, 0, c->stream>>>(p);
    raster_tiles_kernel<0><<<c->sm_count * 4, 256, 0, c->stream>>>(p);
    raster_kernel<1><<<grid, RASTER_THREADS, 0, c->stream>>>(p);
    raster_tiles_kernel<1><<<c->sm_count * 4, 256, 0, c->stream>>>(p);
    resolve_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(p);
    count_launches(5);
    TR_CUDA(cudaGetLastError());
    for (int l = 0; l < 2; l++) {
        c->layer[l].valid = true;
        c->layer[l].has_position = false;
    }
    return TR_OK;
}

}  // namespace tr

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
// Sort-middle, front to back, with a hierarchical Z:
//   A1 bin_count   one thread per triangle: the triangle's 64-triangle chunk against the frame / band planes (bounding sphere, before
//                  any vertex is fetched), exact set-up, cull (back face, off band, w <= 0), count into the 64x64-pixel tile
//                  bins (16 depth buckets per tile and layer), the survivors' 128-byte set-up records
//   A2 bin_scan    exclusive scan of the bin counts
//   A3 bin_fill    scatter the survivors into their bins; its CTA 0 first orders the tile jobs, heaviest first
//   B  raster_tiles  persistent CTAs, one tile job at a time with the tile's depth/id words in shared memory: rounds of 64
//                  triangles (fp32 coarse form with proven bounds + exact double form); the rows of the round's boxes are
//                  handed out by a ticket, one lane solves a row's exact span, one lane per span pixel tests the depth plane,
//                  survivors evaluated exactly, winners by shared-memory atomicMax, 8x8 block minima refreshed per round
//   C  resolve     per pixel: the winning triangle's plane, perspective-correct varyings, the SoA G-buffer planes of both
//                  layers
// (Tried and dropped, round 2: cutting a long tile list into parts that merge by global atomicMax — every part loses the
// others' occluders to the hierarchical Z, the pass got 40 % slower.)
#include <stdlib.h>

#include <algorithm>

#include "tr_internal.h"

using namespace trd;

namespace {

// tile edge in pixels: 64, or 32 when a small band would give too few tiles to fill the GPU (launch_visibility)
constexpr int TILE_THREADS = 256;    // 8 warps per tile CTA
constexpr uint32_t BIN_CAPACITY = 1u << 24;  // (triangle, tile) pairs per frame
constexpr int DEPTH_BUCKETS = 16;    // bin lists per (layer, tile), nearest instances first

struct VisParams {
    const float* positions;
    const float* normals;
    const float* uvs;
    const uint32_t* indices;
    const tr_instance* instances;
    const tr_primitive_info* prims;
    const uint32_t* visible_ids;
    const uint32_t* work_prefix;  // [n_visible + 1]
    const uint32_t* scalars;      // [0] n_visible, [1] total triangles, [6] ~min / [7] max bits of slot_z
    const float* slot_z;          // [n_visible] nearest view depth of the instance in each slot
    const uint4* slot_prim;       // [n_visible] (first_index, draw_buffer_index, first culling chunk, 0) of the slot's primitive, from K1
    const float4* chunk_spheres;  // object-space bounding sphere per TR_CHUNK_TRIS triangles of every primitive; nullptr: no chunk culling
    float4 cull_plane[4];         // world-space planes (xyz, w) whose negative side is off the frame / off the band: left, right, above, below
    float cull_plane_norm[4];     // |xyz| of each, rounded up
    mat4 proj_view;
    float row_y_norm, row_w_norm;  // |rows 1 and 3 of proj_view (xyz)|: how far a unit world offset moves clip y / w
    uint32_t band_cull;            // the band is a strict part of the frame: the work list holds only instances that can reach it
    const uint32_t* list_prefix;   // [n_list + 1] exclusive triangle prefix of the work list (K1's, or the band's own)
    const uint32_t* list_slots;    // [n_list] visible slot of each entry; nullptr = identity
    const uint32_t* list_scalars;  // [0] n_list, [1] triangles in the list
    const uint32_t* list_block_entry;   // [triangles in the list / 256 + 1] the entry that holds triangle 256 b of the list
    uint32_t* band_block_entry;    // the band list's table, written by band_filter_kernel
    uint32_t* band_prefix;         // outputs of band_filter_kernel
    uint32_t* band_slots;
    uint32_t* band_scalars;
    uint32_t width, height, y0, y1;
    unsigned long long* vis[2];
    // sort-middle binning state
    int ts;                                         // tile edge chosen for this launch (64 or 32)
    uint32_t tiles_x, tiles_y, tile_row0, n_tiles;  // tile grid of the band; lists = 2 layers x n_tiles x DEPTH_BUCKETS
    uint32_t n_lists;
    uint32_t* bin_count;          // [n_lists], list = (layer * n_tiles + tile) * DEPTH_BUCKETS + bucket
    uint32_t* bin_start;          // [n_lists + 1]
    uint32_t* bin_cursor;         // [n_lists]
    uint32_t* scan_totals;        // [SCAN_CTAS] slice totals of the scan, zeroed per frame
    uint32_t* bin_entries;        // work-list id of the triangle of each (triangle, list) pair, grouped by list
    uint32_t bin_capacity;
    uint4* records;               // surviving triangles: (slot, tri, tile range, layer)
    struct TriRec* trirec;        // [triangles of the work list] the set-up of a surviving triangle AT ITS WORK-LIST ID, written once
                                  // by pass A1 and read by passes A3, B and C (pass C finds a pixel's winner by the id in its key)
    uint32_t rec_capacity;
    uint32_t* rec_count;
    uint32_t* tile_ticket;
    uint32_t* tile_order;         // [8 * n_tiles] tile jobs, heaviest first (tile_order_kernel)
    uint32_t* n_jobs;             // how many of them
    uint32_t split;               // heavy tiles may be handed out as four quadrant jobs (few tiles for the GPU)
    uint32_t* status;             // bit 0: record overflow, bit 1: bin overflow
    unsigned long long* stats;    // diagnostics: [0] box pixels binned, [1] box pixels after hierarchical Z, [2] exact evaluations
    // resolve outputs
    float* depth[2];
    float* normal[2];
    float* uv[2];
    uint32_t* material_id[2];
    float* scale1;
    const tr_material_info* materials;  // alpha-clip draw buffers only (nullptr: none in the scene)
    const TexDesc* textures;
    uint32_t n_textures;
    float4* duv[2];               // derivative planes (nullptr unless some material binds a texture)
    float2* ddepth[2];
};

struct TriSetup {
    double A[3], B[3], C[3];
    float Z[3], W[3];
    uint32_t vid[3];
    int x_lo, x_hi, y_lo, y_hi;
};

// The set-up of a surviving triangle, 128 bytes, computed ONCE per frame by pass A1 (round 1 recomputed it per (triangle,
// tile) pair in pass B and per (triangle, 32x8 block) in pass C, each time behind a chain of five dependent gathers).
struct __align__(16) TriRec {
    double2 e[4];   // (A0 A1) (A2 B0) (B1 B2) (C0 C1)
    uint4 c2z;      // C2 (two words), Z0, Z1
    float4 z2w;     // Z2, W0, W1, W2
    uint4 ids;      // vid0, vid1, vid2 (set-up order), triangle id in the work list
    uint4 misc;     // instance id, x_lo | x_hi << 16, y_lo | y_hi << 16, layer | depth bucket << 1
};
static_assert(sizeof(TriRec) == 128, "TriRec");

__device__ __forceinline__ void store_trirec(TriRec* r, const TriSetup& s, uint32_t gtid, uint32_t inst_id, uint32_t layer) {
    r->e[0] = make_double2(s.A[0], s.A[1]);
    r->e[1] = make_double2(s.A[2], s.B[0]);
    r->e[2] = make_double2(s.B[1], s.B[2]);
    r->e[3] = make_double2(s.C[0], s.C[1]);
    r->c2z = make_uint4((uint32_t)__double2loint(s.C[2]), (uint32_t)__double2hiint(s.C[2]), __float_as_uint(s.Z[0]), __float_as_uint(s.Z[1]));
    r->z2w = make_float4(s.Z[2], s.W[0], s.W[1], s.W[2]);
    r->ids = make_uint4(s.vid[0], s.vid[1], s.vid[2], gtid);
    r->misc = make_uint4(inst_id, (uint32_t)s.x_lo | ((uint32_t)s.x_hi << 16), (uint32_t)s.y_lo | ((uint32_t)s.y_hi << 16), layer);
}
// edge functions + box (what binning needs)
__device__ __forceinline__ void load_trirec_edges(const TriRec* r, TriSetup& s) {
    const double2 q0 = __ldg(&r->e[0]), q1 = __ldg(&r->e[1]), q2 = __ldg(&r->e[2]), q3 = __ldg(&r->e[3]);
    const uint4 c = __ldg(&r->c2z), m = __ldg(&r->misc);
    s.A[0] = q0.x; s.A[1] = q0.y; s.A[2] = q1.x; s.B[0] = q1.y; s.B[1] = q2.x; s.B[2] = q2.y; s.C[0] = q3.x; s.C[1] = q3.y;
    s.C[2] = __hiloint2double((int)c.y, (int)c.x);
    s.Z[0] = __uint_as_float(c.z); s.Z[1] = __uint_as_float(c.w);
    s.x_lo = (int)(m.y & 0xffffu); s.x_hi = (int)(m.y >> 16); s.y_lo = (int)(m.z & 0xffffu); s.y_hi = (int)(m.z >> 16);
}
__device__ __forceinline__ void load_trirec(const TriRec* r, TriSetup& s, uint32_t& gtid, uint32_t& inst_id) {
    load_trirec_edges(r, s);
    const float4 z = __ldg(&r->z2w);
    const uint4 ids = __ldg(&r->ids);
    s.Z[2] = z.x; s.W[0] = z.y; s.W[1] = z.z; s.W[2] = z.w;
    s.vid[0] = ids.x; s.vid[1] = ids.y; s.vid[2] = ids.z;
    gtid = ids.w;
    inst_id = __ldg(&r->misc.x);
}

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// oracle/raster.c setup_triangle, in three steps (the binning pass queues the triangles between them):
//   setup_front   gather + transform the three vertices, back-face / degenerate test (exact, double)
//   setup_box     Vulkan clip volume, screen bounding box clipped to the band
//   setup_edges   the homogeneous edge functions
struct FrontTri {
    float rx[3], ry[3], Z[3], W[3];   // screen-scaled clip x and y, clip z and w, in set-up order (0, 2, 1)
    uint32_t vid[3];
};

__device__ __forceinline__ bool setup_front(const VisParams& p, const tr_instance* inst, uint32_t first_index, uint32_t tri, FrontTri& f) {
    const float4* iq = reinterpret_cast<const float4*>(inst);
    const float4 ts = __ldg(iq), rot = __ldg(iq + 1);
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
#pragma unroll
    for (int k = 0; k < 3; k++) {
        f.rx[k] = sx[order[k]];
        f.ry[k] = sy[order[k]];
        f.Z[k] = Z[order[k]];
        f.W[k] = W[order[k]];
        f.vid[k] = vid[order[k]];
    }
    return true;
}

__device__ __forceinline__ void setup_edges(const FrontTri& f, TriSetup& s) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        s.A[i] = dsub(dmul(f.ry[a], f.W[b]), dmul(f.W[a], f.ry[b]));
        s.B[i] = dsub(dmul(f.W[a], f.rx[b]), dmul(f.rx[a], f.W[b]));
        s.C[i] = dsub(dmul(f.rx[a], f.ry[b]), dmul(f.ry[a], f.rx[b]));
        s.Z[i] = f.Z[i];
        s.W[i] = f.W[i];
        s.vid[i] = f.vid[i];
    }
}

__device__ __forceinline__ bool setup_box(const VisParams& p, const FrontTri& s, int& bx_lo, int& bx_hi, int& by_lo, int& by_hi) {
    const float* rx = s.rx;
    const float* ry = s.ry;
    // Vulkan clip volume: w > 0 and z <= w.  Every vertex at/behind the camera plane => no fragment.
    if (!(s.W[0] > 0.0f) && !(s.W[1] > 0.0f) && !(s.W[2] > 0.0f)) return false;
    int x_lo = 0, x_hi = (int)p.width - 1, y_lo = (int)p.y0, y_hi = (int)p.y1 - 1;
    const float inf = __int_as_float(0x7f800000);
    float mnx = inf, mxx = -inf, mny = inf, mxy = -inf, pad = 0.0f;
    bool whole = false;
    if (s.W[0] > 0.0f && s.W[1] > 0.0f && s.W[2] > 0.0f) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float px = xdiv(rx[k], s.W[k]), py = xdiv(ry[k], s.W[k]);
            mnx = rmin(mnx, px); mxx = rmax(mxx, px);
            mny = rmin(mny, py); mxy = rmax(mxy, py);
        }
    } else {
        // crosses the camera plane: bound the part inside the near plane (z <= w) — oracle/raster.c setup_triangle
        float nd[3];
        int inside = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            nd[k] = xsub(s.W[k], s.Z[k]);
            if (nd[k] >= 0.0f) inside++;
        }
        if (inside == 0) return false;
        pad = 1.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int j = (k + 1) % 3;
            if (nd[k] >= 0.0f) {
                if (!(s.W[k] > 0.0f)) whole = true;
                else {
                    const float px = xdiv(rx[k], s.W[k]), py = xdiv(ry[k], s.W[k]);
                    mnx = rmin(mnx, px); mxx = rmax(mxx, px);
                    mny = rmin(mny, py); mxy = rmax(mxy, py);
                }
            }
            if ((nd[k] >= 0.0f) != (nd[j] >= 0.0f)) {
                const bool k_in = nd[k] >= 0.0f;
                const float na = k_in ? nd[k] : nd[j], nb = k_in ? nd[j] : nd[k];
                const float rxa = k_in ? rx[k] : rx[j], rxb = k_in ? rx[j] : rx[k];
                const float rya = k_in ? ry[k] : ry[j], ryb = k_in ? ry[j] : ry[k];
                const float wa = k_in ? s.W[k] : s.W[j], wb = k_in ? s.W[j] : s.W[k];
                const float t = xdiv(na, xsub(na, nb));
                const float cx = xadd(rxa, xmul(t, xsub(rxb, rxa)));
                const float cy = xadd(rya, xmul(t, xsub(ryb, rya)));
                const float cw = xadd(wa, xmul(t, xsub(wb, wa)));
                if (!(cw > 0.0f)) whole = true;
                else {
                    const float px = xdiv(cx, cw), py = xdiv(cy, cw);
                    mnx = rmin(mnx, px); mxx = rmax(mxx, px);
                    mny = rmin(mny, py); mxy = rmax(mxy, py);
                }
            }
        }
    }
    if (!whole) {
        mnx = xsub(mnx, pad); mny = xsub(mny, pad); mxx = xadd(mxx, pad); mxy = xadd(mxy, pad);
        if (!(mxx >= 0.0f) || !(mnx <= (float)p.width) || !(mxy >= (float)p.y0) || !(mny <= (float)p.y1)) return false;
        const float fx_lo = floorf(xsub(mnx, 0.5f)), fx_hi = ceilf(xsub(mxx, 0.5f));
        const float fy_lo = floorf(xsub(mny, 0.5f)), fy_hi = ceilf(xsub(mxy, 0.5f));
        if (fx_lo > (float)x_lo) x_lo = (int)fx_lo;
        if (fx_hi < (float)x_hi) x_hi = (int)fx_hi;
        if (fy_lo > (float)y_lo) y_lo = (int)fy_lo;
        if (fy_hi < (float)y_hi) y_hi = (int)fy_hi;
    }
    if (x_lo > x_hi || y_lo > y_hi) return false;
    bx_lo = x_lo; bx_hi = x_hi; by_lo = y_lo; by_hi = y_hi;
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
    const double r = __ddiv_rn(1.0, S);
    l[0] = __double2float_rn(dmul(E[0], r));
    l[1] = __double2float_rn(dmul(E[1], r));
    l[2] = __double2float_rn(dmul(E[2], r));
    const float zq = xadd(xadd(xmul(l[0], s.Z[0]), xmul(l[1], s.Z[1])), xmul(l[2], s.Z[2]));
    const float wq = xadd(xadd(xmul(l[0], s.W[0]), xmul(l[1], s.W[1])), xmul(l[2], s.W[2]));
    if (!(wq > 0.0f)) return false;
    const float d = xdiv(zq, wq);
    if (!(d > 0.0f) || d > 1.0f) return false;
    depth = d;
    return true;
}

// oracle/raster.c eval_plane: barycentrics / depth of the triangle's plane at a pixel centre without the coverage test
__device__ __forceinline__ bool eval_plane(const TriSetup& s, int px, int py, float l[3], float& depth) {
    const double qx = (double)px + 0.5, qy = (double)py + 0.5;
    double E[3];
#pragma unroll
    for (int i = 0; i < 3; i++) E[i] = dadd(dadd(dmul(s.A[i], qx), dmul(s.B[i], qy)), s.C[i]);
    const double S = dadd(dadd(E[0], E[1]), E[2]);
    if (!(S > 0.0)) return false;
    const double r = __ddiv_rn(1.0, S);
    l[0] = __double2float_rn(dmul(E[0], r));
    l[1] = __double2float_rn(dmul(E[1], r));
    l[2] = __double2float_rn(dmul(E[2], r));
    const float zq = xadd(xadd(xmul(l[0], s.Z[0]), xmul(l[1], s.Z[1])), xmul(l[2], s.Z[2]));
    const float wq = xadd(xadd(xmul(l[0], s.W[0]), xmul(l[1], s.W[1])), xmul(l[2], s.W[2]));
    if (!(wq > 0.0f)) return false;
    depth = xdiv(zq, wq);
    return true;
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

__device__ __forceinline__ uint32_t find_slot(const VisParams& p, uint32_t w, uint32_t n_visible) {
    uint32_t lo = 0, hi = n_visible;  // largest slot with work_prefix[slot] <= w
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(p.work_prefix + mid) <= w) lo = mid; else hi = mid;
    }
    return lo;
}

// The bin lists a triangle goes to: every tile of its bounding box, edge-tested when the box spans more than
// 4 tiles.  Called by all 32 lanes of a warp together: boxes of up to 16 tiles are walked by their own lane,
// larger ones (a triangle crossing the camera plane spans the whole band) by the whole warp, one tile per
// lane, so that no single thread ever walks thousands of tiles.
// f(list, payload) handles one (triangle, list) pair; g(list, payload, peers) handles the lanes of a warp whose
// single-tile triangles all go to the same list (the common case: neighbouring triangles of one mesh), so that they
// can share one atomic.
template <typename F, typename G>
__device__ __forceinline__ void bin_triangle(const VisParams& p, bool keep, const TriSetup& s, uint32_t range, uint32_t layer,
                                             uint2 payload, F f, G g) {
    const uint32_t lane = threadIdx.x & 31;
    const int tx0 = range & 0xff, tx1 = (range >> 8) & 0xff, ty0 = (range >> 16) & 0xff, ty1 = range >> 24;
    const int ntx = tx1 - tx0 + 1, n_tiles = ntx * (ty1 - ty0 + 1);
    auto visit = [&](const TriSetup& t, int tx, int ty, bool test, uint32_t lay, uint2 pl) {
        if (test) {
            const int x0 = max(tx * p.ts, t.x_lo), x1 = min(tx * p.ts + p.ts - 1, t.x_hi);
            const int y0 = max((ty + (int)p.tile_row0) * p.ts, t.y_lo), y1 = min((ty + (int)p.tile_row0) * p.ts + p.ts - 1, t.y_hi);
            if (!tile_may_overlap(t, x0, y0, x1, y1)) return;
        }
        f((((lay & 1u) * p.n_tiles + (uint32_t)ty * p.tiles_x + (uint32_t)tx) * DEPTH_BUCKETS) + (lay >> 1), pl);
    };
    const bool single = keep && n_tiles == 1;
    const uint32_t single_mask = __ballot_sync(0xffffffffu, single);
    if (single) {
        const uint32_t list = (((layer & 1u) * p.n_tiles + (uint32_t)ty0 * p.tiles_x + (uint32_t)tx0) * DEPTH_BUCKETS) + (layer >> 1);
        g(list, payload, __match_any_sync(single_mask, list));
    }
    if (keep && n_tiles > 1 && n_tiles <= 16)
        for (int ty = ty0; ty <= ty1; ty++)
            for (int tx = tx0; tx <= tx1; tx++) visit(s, tx, ty, n_tiles > 4, layer, payload);
    uint32_t big = __ballot_sync(0xffffffffu, keep && n_tiles > 16);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        TriSetup t;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            t.A[i] = __shfl_sync(0xffffffffu, s.A[i], src);
            t.B[i] = __shfl_sync(0xffffffffu, s.B[i], src);
            t.C[i] = __shfl_sync(0xffffffffu, s.C[i], src);
        }
        t.x_lo = __shfl_sync(0xffffffffu, s.x_lo, src);
        t.x_hi = __shfl_sync(0xffffffffu, s.x_hi, src);
        t.y_lo = __shfl_sync(0xffffffffu, s.y_lo, src);
        t.y_hi = __shfl_sync(0xffffffffu, s.y_hi, src);
        const uint32_t r = __shfl_sync(0xffffffffu, range, src), lay = __shfl_sync(0xffffffffu, layer, src);
        const uint2 pl = make_uint2(__shfl_sync(0xffffffffu, payload.x, src), __shfl_sync(0xffffffffu, payload.y, src));
        const int bx0 = r & 0xff, bx1 = (r >> 8) & 0xff, by0 = (r >> 16) & 0xff, by1 = r >> 24;
        const int bnx = bx1 - bx0 + 1, bn = bnx * (by1 - by0 + 1);
        for (int i = (int)lane; i < bn; i += 32) visit(t, bx0 + i % bnx, by0 + i / bnx, true, lay, pl);
    }
}

// front-to-back bucket of an instance: log-spaced in its nearest view depth over the visible set's range
__device__ __forceinline__ uint32_t depth_bucket(const VisParams& p, uint32_t slot) {
    const float z_hi = __uint_as_float(p.scalars[7]);
    const float z_lo = fmaxf(__uint_as_float(~p.scalars[6]), z_hi * (1.0f / 64.0f));  // at most 6 octaves: instances that
    if (!(z_hi > z_lo)) return 0u;                                                     // straddle the camera share bucket 0
    const float t = __log2f(__ldg(p.slot_z + slot) / z_lo) / __log2f(z_hi / z_lo) * (float)DEPTH_BUCKETS;
    return (uint32_t)min(max((int)t, 0), DEPTH_BUCKETS - 1);
}

// slot of the first lane's work item by binary search, the other lanes walk forward from it (they are
// almost always in the same or the next instance)
__device__ __forceinline__ uint32_t find_entry_warp(const uint32_t* prefix, uint32_t base, uint32_t w, uint32_t n) {
    uint32_t e = 0;
    if ((threadIdx.x & 31) == 0) {
        uint32_t lo = 0, hi = n;  // largest entry with prefix[entry] <= base
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(prefix + mid) <= base) lo = mid; else hi = mid;
        }
        e = lo;
    }
    e = __shfl_sync(0xffffffffu, e, 0);
    while (e + 1 < n && __ldg(prefix + e + 1) <= w) e++;
    return e;
}

// conservative rows of an instance's bounding sphere: clip(C + d) = clip(C) + M d, |d| <= r
__device__ __forceinline__ bool instance_on_band(const VisParams& p, const tr_instance* inst, const tr_primitive_info* prim) {
    const float4 sph = __ldg(reinterpret_cast<const float4*>(prim));
    const float4 ts = __ldg(reinterpret_cast<const float4*>(inst)), rot = __ldg(reinterpret_cast<const float4*>(inst) + 1);
    const f3 c = xadd3(mk3(ts.x, ts.y, ts.z), xscale3(xquat_mul3(rot.x, rot.y, rot.z, rot.w, mk3(sph.x, sph.y, sph.z)), ts.w));
    const f4 cc = xmat4_mul(p.proj_view, c.x, c.y, c.z, 1.0f);
    const float r = fabsf(sph.w * ts.w) * 1.01f + 1e-3f;
    const float w_lo = cc.w - r * p.row_w_norm, w_hi = cc.w + r * p.row_w_norm;
    if (!(w_lo > 0.0f)) return true;
    const float y_lo = cc.y - r * p.row_y_norm, y_hi = cc.y + r * p.row_y_norm;
    const float n_lo = fminf(y_lo / w_lo, y_lo / w_hi), n_hi = fmaxf(y_hi / w_lo, y_hi / w_hi);
    const float half_h = 0.5f * (float)p.height;
    return (n_hi + 1.0f) * half_h + 2.0f >= (float)p.y0 && (n_lo + 1.0f) * half_h - 2.0f <= (float)p.y1;
}

// ---- pass A0 (band-sharded frames only): the band's own work list — visible instances whose bounding sphere can reach
// the band's rows, with their exclusive triangle prefix — so that pass A1 does not walk the whole scene's triangles on
// every rank.  One CTA (the visible set is a few thousand instances).
__global__ void __launch_bounds__(1024) band_filter_kernel(const __grid_constant__ VisParams p) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const uint32_t n_visible = p.scalars[0], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    constexpr uint32_t PER = 4;   // consecutive slots per thread: their gathers are in flight together
    for (uint32_t chunk = 0; chunk < n_visible; chunk += 1024 * PER) {
        unsigned long long v[PER], mine = 0;
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) {
            const uint32_t slot = chunk + tid * PER + k;
            bool on = false;
            uint32_t tris = 0;
            if (slot < n_visible) {
                const tr_instance* inst = p.instances + __ldg(p.visible_ids + slot);
                const tr_primitive_info* prim = p.prims + __ldg(&inst->primitive_id);
                on = instance_on_band(p, inst, prim);
                if (on) tris = __ldg(p.work_prefix + slot + 1) - __ldg(p.work_prefix + slot);
            }
            v[k] = ((unsigned long long)tris << 24) | (on ? 1ull : 0ull);
            mine += v[k];
        }
        unsigned long long incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long off = s_carry, tot = 0;
        for (uint32_t k = 0; k < 32; k++) {
            if (k < warp) off += s_warp[k];
            tot += s_warp[k];
        }
        unsigned long long excl = off + incl - mine;
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) {
            if (v[k] & 1ull) {
                const uint32_t at = (uint32_t)(excl & 0xffffffull);
                p.band_slots[at] = chunk + tid * PER + k;
                const uint32_t t0 = (uint32_t)(excl >> 24), tris = (uint32_t)(v[k] >> 24);
                p.band_prefix[at] = t0;
                for (uint32_t b = (t0 + 255u) >> 8; (b << 8) < t0 + tris; b++) p.band_block_entry[b] = at;
            }
            excl += v[k];
        }
        __syncthreads();
        if (tid == 0) s_carry += tot;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t n = (uint32_t)(s_carry & 0xffffffull), t = (uint32_t)(s_carry >> 24);
        p.band_scalars[0] = n;
        p.band_scalars[1] = t;
        p.band_prefix[n] = t;
    }
}

// ---- pass A1: set up every triangle of the visible instances once, keep the survivors (front-facing,
// on-screen, inside the band) as compact records and count them into the per-tile bin lists.
#ifndef TR_BIN_CTAS
#define TR_BIN_CTAS 3
#endif
#ifndef TR_BIN_WAVES   // CTAs of the binning pass per resident slot: their ranges cost very unequal amounts (survivors)
#define TR_BIN_WAVES 8
#endif
// Two steps with a queue between them: only about one triangle in seven survives the back-face and box tests, and a warp
// that carried its few survivors through the edge set-up, the binning and the 128-byte record store lane by lane would run
// that half of the pass at a seventh of its lanes.  Survivors wait in the warp's shared-memory queue until 32 are there.
constexpr int BINQ_WORDS = 20, BINQ_CAP = 64;   // per entry: rx ry Z W (12), vid (3), slot, tri, layer, box x, box y
__global__ void __launch_bounds__(256, TR_BIN_CTAS) bin_count_kernel(const __grid_constant__ VisParams p) {
    __shared__ uint32_t s_queue[8][BINQ_WORDS][BINQ_CAP];
    const uint32_t n_list = p.list_scalars[0], total = p.list_scalars[1];
    const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
    uint32_t (*q)[BINQ_CAP] = s_queue[threadIdx.x >> 5];
    // a CTA takes a contiguous range of the work list, so that (i) a warp's next 32 triangles are 256 further on and the list
    // entry is found by walking on from the last one, (ii) the records of neighbouring triangles lie together in memory —
    // a tile's bin list is mostly runs of neighbouring triangles, and pass B gathers their records
    const uint32_t per = ((total + gridDim.x - 1) / gridDim.x + 255u) & ~255u;
    const uint32_t range_begin = blockIdx.x * per, range_end = min(range_begin + per, total);
    uint32_t entry = 0, qn = 0;
    bool have_entry = false;

    // step 2 on the queue entries [from, from + 32) (fewer at the very end): edge functions, bin counts, the record
    auto finish = [&](uint32_t from, uint32_t n) {
        const bool keep = lane < n;
        const uint32_t e = from + min(lane, n - 1u);
        FrontTri f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            f.rx[k] = __uint_as_float(q[k][e]);
            f.ry[k] = __uint_as_float(q[3 + k][e]);
            f.Z[k] = __uint_as_float(q[6 + k][e]);
            f.W[k] = __uint_as_float(q[9 + k][e]);
            f.vid[k] = q[12 + k][e];
        }
        const uint32_t slot = q[15][e], tri = q[16][e], layer = q[17][e], bx = q[18][e], by = q[19][e];
        TriSetup s;
        setup_edges(f, s);
        s.x_lo = (int)(bx & 0xffffu); s.x_hi = (int)(bx >> 16); s.y_lo = (int)(by & 0xffffu); s.y_hi = (int)(by >> 16);
        const uint32_t range = (uint32_t)(s.x_lo / p.ts) | ((uint32_t)(s.x_hi / p.ts) << 8) |
                               ((uint32_t)(s.y_lo / p.ts - (int)p.tile_row0) << 16) | ((uint32_t)(s.y_hi / p.ts - (int)p.tile_row0) << 24);
        bin_triangle(p, keep, s, range, layer, make_uint2(0, 0), [&](uint32_t list, uint2) { atomicAdd(p.bin_count + list, 1u); },
                     [&](uint32_t list, uint2, uint32_t peers) {
                         if (lane == (uint32_t)__ffs(peers) - 1u) atomicAdd(p.bin_count + list, (uint32_t)__popc(peers));
                     });
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(p.rec_count, n);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (keep) {
            const uint32_t r = first + lane;
            const uint32_t gtid = __ldg(p.work_prefix + slot) + tri;
            if (r < p.rec_capacity && gtid < p.rec_capacity) {
                p.records[r] = make_uint4(gtid, 0u, range, layer);
                store_trirec(p.trirec + gtid, s, gtid, __ldg(p.visible_ids + slot), layer);
            } else atomicOr(p.status, 1u);
        }
    };

    for (uint32_t base = range_begin + (threadIdx.x & ~31u); base < range_end; base += 256u) {
        const uint32_t w = base + lane;
        {   // the entry of the range's first triangle comes from K1's (or the band filter's) table; from there, and from the
            // warp's previous triangles 256 before these, it is a short walk
            if (!have_entry) {
                entry = p.list_block_entry ? __ldg(p.list_block_entry + (base >> 8)) : find_entry_warp(p.list_prefix, base, min(w, total - 1), n_list);
                have_entry = true;
            }
            const uint32_t wc = min(w, total - 1);
            while (entry + 1 < n_list && __ldg(p.list_prefix + entry + 1) <= wc) entry++;
        }
        // step 1: vertices, back face, clip volume and box
        bool keep = false;
        FrontTri f;
        uint32_t slot = 0, tri = 0, layer = 0;
        int x_lo = 0, x_hi = 0, y_lo = 0, y_hi = 0;
        if (w < total) {
            slot = p.list_slots ? __ldg(p.list_slots + entry) : entry;
            tri = w - __ldg(p.list_prefix + entry);
            const tr_instance* inst = p.instances + __ldg(p.visible_ids + slot);
            const uint4 prim = __ldg(p.slot_prim + slot);
            const uint32_t bucket = prim.y;
            layer = bucket >> 1;  // draw buffers 0/1 (opaque, alpha clip) -> layer 0, 2/3 -> the transmissive layer
            // The chunk of 64 triangles this one belongs to (the warp's 32 share it, so the branch is uniform): when its
            // bounding sphere lies wholly on the outer side of a frame / band plane, no pixel of the band can see it.  Half of
            // the triangles of the VISIBLE instances of the 10 k-instance scene go this way, before any vertex is fetched.
            bool chunk_on = true;
            if (p.chunk_spheres) {
                const float4 sph = __ldg(p.chunk_spheres + prim.z + tri / TR_CHUNK_TRIS);
                const float4 ts = __ldg(reinterpret_cast<const float4*>(inst)), rot = __ldg(reinterpret_cast<const float4*>(inst) + 1);
                const f3 wc = xadd3(mk3(ts.x, ts.y, ts.z), xscale3(xquat_mul3(rot.x, rot.y, rot.z, rot.w, mk3(sph.x, sph.y, sph.z)), ts.w));
                const float r = fabsf(sph.w * ts.w) * 1.01f + 1e-3f;   // generous against the fp32 rounding of the centre and the planes
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 pl = p.cull_plane[q];
                    const float d = fmaf(pl.x, wc.x, fmaf(pl.y, wc.y, fmaf(pl.z, wc.z, pl.w)));
                    const float slack = r * p.cull_plane_norm[q] + 1e-4f * (fabsf(pl.x * wc.x) + fabsf(pl.y * wc.y) + fabsf(pl.z * wc.z) + fabsf(pl.w));
                    if (d < -slack) chunk_on = false;
                }
            }
            keep = chunk_on && bucket < 4u && setup_front(p, inst, prim.x, tri, f) && setup_box(p, f, x_lo, x_hi, y_lo, y_hi);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const uint32_t e = qn + (uint32_t)__popc(mask & lt_mask);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                q[k][e] = __float_as_uint(f.rx[k]);
                q[3 + k][e] = __float_as_uint(f.ry[k]);
                q[6 + k][e] = __float_as_uint(f.Z[k]);
                q[9 + k][e] = __float_as_uint(f.W[k]);
                q[12 + k][e] = f.vid[k];
            }
            q[15][e] = slot;
            q[16][e] = tri;
            q[17][e] = layer | (depth_bucket(p, slot) << 1);   // layer | depth bucket << 1 travels with the record
            q[18][e] = (uint32_t)x_lo | ((uint32_t)x_hi << 16);
            q[19][e] = (uint32_t)y_lo | ((uint32_t)y_hi << 16);
        }
        qn += (uint32_t)__popc(mask);
        __syncwarp();
        if (qn >= 32u) {
            qn -= 32u;
            finish(qn, 32u);
            __syncwarp();
        }
    }
    if (qn) finish(0u, qn);
}

// ---- pass A2: exclusive scan of the bin counts.  SCAN_CTAS co-resident CTAs (far fewer than SMs) each scan a contiguous
// slice, publish its total, and add the totals of the slices before it (which they wait for; all CTAs are running, so
// the wait cannot deadlock).
constexpr int SCAN_CTAS = 32;
__global__ void __launch_bounds__(1024) bin_scan_kernel(const __grid_constant__ VisParams p) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry, s_prev;
    const uint32_t n = p.n_lists, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;  // n is a multiple of 32
    const uint32_t per = ((n / 4 + gridDim.x - 1) / gridDim.x) * 4;                      // lists per CTA, a multiple of 4
    const uint32_t begin = min(blockIdx.x * per, n), end = min(begin + per, n);
    // pass 1: this slice's total
    uint32_t sum = 0;
    for (uint32_t i = begin + tid * 4; i < end; i += 4096) {
        const uint4 q = *reinterpret_cast<const uint4*>(p.bin_count + i);
        sum += q.x + q.y + q.z + q.w;
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
        for (int k = 0; k < 32; k++) t += s_warp[k];
        // flag in the top bit: totals are < 2^31 (bin_capacity is 2^24 and larger sums only set the overflow bit)
        atomicExch(p.scan_totals + blockIdx.x, 0x80000000u | min(t, 0x7fffffffu));
        uint32_t prev = 0;
        for (uint32_t b = 0; b < blockIdx.x; b++) {
            uint32_t v;
            do { v = *reinterpret_cast<volatile uint32_t*>(p.scan_totals + b); } while (!(v & 0x80000000u));
            prev += v & 0x7fffffffu;
        }
        s_prev = prev;
        s_carry = 0;
    }
    __syncthreads();
    // pass 2: exclusive scan of the slice on top of s_prev
    for (uint32_t chunk = begin; chunk < end; chunk += 4096) {
        const uint32_t i0 = chunk + tid * 4;
        const uint4 q = i0 < end ? *reinterpret_cast<const uint4*>(p.bin_count + i0) : make_uint4(0, 0, 0, 0);
        const uint32_t s4 = q.x + q.y + q.z + q.w;
        uint32_t incl = s4;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        __syncthreads();
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t off = s_prev + s_carry, tot = 0;
#pragma unroll
        for (uint32_t k = 0; k < 32; k++) {
            const uint32_t t = s_warp[k];
            if (k < warp) off += t;
            tot += t;
        }
        if (i0 < end) {
            uint4 o;
            o.x = off + incl - s4;
            o.y = o.x + q.x;
            o.z = o.y + q.y;
            o.w = o.z + q.z;
            *reinterpret_cast<uint4*>(p.bin_start + i0) = o;
            *reinterpret_cast<uint4*>(p.bin_cursor + i0) = o;
        }
        __syncthreads();
        if (tid == 0) s_carry += tot;
        __syncthreads();
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        const uint32_t total = s_prev + s_carry;
        p.bin_start[n] = total;
        if (total > p.bin_capacity) atomicOr(p.status, 2u);
    }
}

// ---- between A2 and B: the order in which the tile jobs are handed out.  The tile kernel's CTAs take jobs from a ticket
// counter; a heavy tile taken late is the tail of the kernel, so the jobs are sorted by their triangle count, heaviest first
// (counting sort on the bit length of the count — longest-processing-time-first needs no finer order), empty tiles last.
// When the band has few tiles for the GPU (a rank of a many-GPU frame: a 270-row band of a 4K frame is 600 jobs for 592 CTA
// slots, and the pass lasts as long as its heaviest tile), a tile that holds well over the average is handed out as four
// 32x32-pixel quadrant jobs: each walks the tile's whole list but keeps only what reaches its quadrant.  (Cutting the LIST
// instead was tried in round 2 and lost 40 %: every part loses the other parts' occluders to the hierarchical Z.)
// A job word: item | (1 + quadrant) << 28, or the bare item for a whole tile.
// Run by CTA 0 of the fill pass (pass A3) before its share of the filling: one launch fewer, and it is off the critical path.
#ifndef TR_SPLIT_NUM   // a tile is split when it holds more than TR_SPLIT_NUM / TR_SPLIT_DEN of the mean
#define TR_SPLIT_NUM 3
#define TR_SPLIT_DEN 2
#endif
__device__ __forceinline__ void tile_order_block(const VisParams& p) {
    __shared__ uint32_t s_hist[33], s_base[33];
    const uint32_t tid = threadIdx.x, n = 2u * p.n_tiles, nt = blockDim.x;
    if (tid < 33) s_hist[tid] = 0;
    __syncthreads();
    const uint32_t split_above = p.split ? max(128u, (uint32_t)(((unsigned long long)min(p.bin_start[p.n_lists], p.bin_capacity) * (unsigned long long)TR_SPLIT_NUM) / ((unsigned long long)TR_SPLIT_DEN * n))) : 0xffffffffu;
    auto weight = [&](uint32_t item, bool& split) {
        const uint32_t begin = min(p.bin_start[item * DEPTH_BUCKETS], p.bin_capacity), end = min(p.bin_start[(item + 1) * DEPTH_BUCKETS], p.bin_capacity);
        split = end - begin > split_above;
        return split ? (end - begin) / 4u + 1u : end - begin;
    };
    auto bucket = [](uint32_t w) { return 32u - (uint32_t)__clz(w); };   // 0 for an empty tile
    for (uint32_t i = tid; i < n; i += nt) {
        bool split;
        const uint32_t w = weight(i, split);
        atomicAdd(&s_hist[bucket(w)], split ? 4u : 1u);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t acc = 0;
        for (int b = 32; b >= 0; b--) {
            s_base[b] = acc;
            acc += s_hist[b];
        }
        *p.n_jobs = acc;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += nt) {
        bool split;
        const uint32_t w = weight(i, split);
        if (split) {
            const uint32_t at = atomicAdd(&s_base[bucket(w)], 4u);
            for (uint32_t q = 0; q < 4; q++) p.tile_order[at + q] = i | ((1u + q) << 28);
        } else {
            p.tile_order[atomicAdd(&s_base[bucket(w)], 1u)] = i;
        }
    }
}

// ---- pass A3: scatter the surviving triangles into their bin lists
__global__ void __launch_bounds__(256, TR_BIN_CTAS) bin_fill_kernel(const __grid_constant__ VisParams p) {
    if (blockIdx.x == 0) tile_order_block(p);
    const uint32_t n = min(*p.rec_count, p.rec_capacity);
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t r = base + lane;
        const bool keep = r < n;
        uint4 rec = make_uint4(0, 0, 0, 0);
        TriSetup s{};
        if (keep) {
            rec = p.records[r];
            const int tx0 = rec.z & 0xff, tx1 = (rec.z >> 8) & 0xff, ty0 = (rec.z >> 16) & 0xff, ty1 = rec.z >> 24;
            if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 4) load_trirec_edges(p.trirec + rec.x, s);  // the edge test of large triangles
        }
        bin_triangle(p, keep, s, rec.z, rec.w, make_uint2(rec.x, 0u), [&](uint32_t list, uint2 pl) {
            const uint32_t pos = atomicAdd(p.bin_cursor + list, 1u);
            if (pos < p.bin_capacity) p.bin_entries[pos] = pl.x;
        }, [&](uint32_t list, uint2 pl, uint32_t peers) {
            const int leader = __ffs(peers) - 1;
            uint32_t first = 0;
            if ((int)lane == leader) first = atomicAdd(p.bin_cursor + list, (uint32_t)__popc(peers));
            first = __shfl_sync(peers, first, leader);
            const uint32_t pos = first + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            if (pos < p.bin_capacity) p.bin_entries[pos] = pl.x;
        });
    }
}

// ---- pass B: persistent CTAs take (layer, tile) jobs from a ticket; the job's depth/id words live in shared memory.
// The CTA takes ROUND (64) triangles of the bin list per round: thread i loads record i and parks an fp32 "coarse" form
// and the exact double form in shared memory while the upper half of the CTA refreshes the 8x8-block depth minima; a
// triangle whose nearest depth lies behind every block its box touches is dropped.  The rows of the round's boxes are laid
// end to end and handed to the warps 32 at a time: one lane per row solves the three fp32 edge tests (rigorous error
// bound) for their exact interval, then one lane per span pixel tests a conservative fp32 depth plane against the tile's
// current depth; the survivors are queued per warp and evaluated 32 at a time with the exact double rule (eval_pixel), so
// the expensive path runs with full warps.
#ifndef TR_ROUND
#define TR_ROUND 64
#endif
#ifndef TR_PREFETCH
#define TR_PREFETCH 1
#endif
#ifndef TR_TAIL_ROWS
#define TR_TAIL_ROWS 256
#endif
constexpr int ROUND = TR_ROUND;  // triangles per round (<= TILE_THREADS); smaller rounds refresh the hierarchical Z more often
struct TileRecs {
    double A[3][ROUND], B[3][ROUND], C[3][ROUND];
    float ea[3][ROUND], eb[3][ROUND], ec[3][ROUND], ebound[3][ROUND];
    float d0[ROUND], gx[ROUND], gy[ROUND], margin[ROUND];
    float Z[3][ROUND], W[3][ROUND];
    uint32_t gtid[ROUND];
    uint32_t clip_mat[ROUND];  // material id of an alpha-clip triangle, 0xffffffff otherwise
    uint2 entry[ROUND];        // (slot, tri): the alpha test re-reads the triangle's uvs
    uint32_t box[ROUND];   // x_lo | y_lo << 6 | (bw - 1) << 12
    uint32_t off[ROUND + 1];   // exclusive prefix of the clipped boxes' heights: the rows of a round, end to end
    uint4 span[TILE_THREADS / 32][32];     // per warp: the spans of 32 rows (first sample, first pixel's queue word, depth plane)
    float span_mg[TILE_THREADS / 32][32];
    uint32_t queue[TILE_THREADS / 32][64];
    uint32_t warp_tot[TILE_THREADS / 32];
    uint32_t row_ticket;
    __align__(16) float zmin_blk[64];    // hierarchical Z: min depth of each 8x8 pixel block, refreshed after every round
};

// depth_pre_pass_alpha_clip (shader/src/lib.rs:269-293): diffuse alpha (factor x texture, implicit level of detail from
// the uv differences to the right / lower neighbour on the triangle's plane) against the material's cutoff
__device__ __noinline__ bool alpha_test_kills(const VisParams& p, const TriSetup& s, const float l[3], uint2 entry, uint32_t mat_id,
                                               int px, int py) {
    const tr_material_info* m = p.materials + mat_id;
    float alpha = __ldg(&m->diffuse_factor.w);
    const int32_t t_diffuse = __ldg(&m->textures.diffuse);
    if (t_diffuse != -1) {
        const tr_instance* inst = p.instances + __ldg(p.visible_ids + entry.x);
        const tr_primitive_info* prim = p.prims + __ldg(&inst->primitive_id);
        const uint32_t base = __ldg(&prim->first_index) + entry.y * 3;
        const uint32_t v0 = __ldg(p.indices + base), v1 = __ldg(p.indices + base + 2), v2 = __ldg(p.indices + base + 1);  // setup order (0, 2, 1)
        const float *u0 = p.uvs + (size_t)v0 * 2, *u1 = p.uvs + (size_t)v1 * 2, *u2 = p.uvs + (size_t)v2 * 2;
        const float uu = xadd(xadd(xmul(l[0], __ldg(u0)), xmul(l[1], __ldg(u1))), xmul(l[2], __ldg(u2)));
        const float vv = xadd(xadd(xmul(l[0], __ldg(u0 + 1)), xmul(l[1], __ldg(u1 + 1))), xmul(l[2], __ldg(u2 + 1)));
        float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            float ln[3], dn;
            if (eval_plane(s, px + (k == 0), py + (k == 1), ln, dn)) {
                dq[k * 2] = xsub(xadd(xadd(xmul(ln[0], __ldg(u0)), xmul(ln[1], __ldg(u1))), xmul(ln[2], __ldg(u2))), uu);
                dq[k * 2 + 1] = xsub(xadd(xadd(xmul(ln[0], __ldg(u0 + 1)), xmul(ln[1], __ldg(u1 + 1))), xmul(ln[2], __ldg(u2 + 1))), vv);
            }
        }
        float a = 0.0f;
        if (p.textures && (uint32_t)t_diffuse < p.n_textures && p.textures[t_diffuse].base)
            a = sample_texture(p.textures[t_diffuse], uu, vv, make_float4(dq[0], dq[1], dq[2], dq[3])).w;
        alpha = xmul(alpha, a);
    }
    return alpha < __ldg(&m->alpha_clipping_cutoff);
}

template <int TS, bool CLIP>
__device__ __forceinline__ void exact_sample(const VisParams& p, const TileRecs& tr_, unsigned long long* keys, uint32_t q, int tile_x0,
                                             int tile_y0) {
    const uint32_t j = q & 255u, lx = (q >> 8) & 63u, ly = (q >> 14) & 63u;
    TriSetup s;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        s.A[i] = tr_.A[i][j];
        s.B[i] = tr_.B[i][j];
        s.C[i] = tr_.C[i][j];
        s.Z[i] = tr_.Z[i][j];
        s.W[i] = tr_.W[i][j];
    }
    float l[3], d;
    if (!eval_pixel(s, tile_x0 + (int)lx, tile_y0 + (int)ly, l, d)) return;
    const uint32_t clip_mat = CLIP ? tr_.clip_mat[j] : 0xffffffffu;
    if (CLIP && clip_mat != 0xffffffffu && alpha_test_kills(p, s, l, tr_.entry[j], clip_mat, tile_x0 + (int)lx, tile_y0 + (int)ly)) return;
    const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(0xffffffffu - tr_.gtid[j]);
    unsigned long long* k = keys + ly * TS + lx;
    if (*reinterpret_cast<volatile unsigned long long*>(k) < key) atomicMax(k, key);
}

#ifndef TR_TILE_CTAS
#define TR_TILE_CTAS 4
#endif
template <int TS, bool CLIP>
__global__ void __launch_bounds__(TILE_THREADS, CLIP ? 2 : TR_TILE_CTAS) raster_tiles_kernel(const __grid_constant__ VisParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    TileRecs& R = *reinterpret_cast<TileRecs*>(smem_raw + TS * TS * 8);
    __shared__ uint32_t s_item;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* queue = R.queue[warp];
    // 32-bit shared-window addresses of the hot structures, made opaque to the optimiser so that they live in registers:
    // otherwise every access in the sample loop re-derives its address (S2R + LEA + ...), which costs more than the loads
    uint32_t recs_s = (uint32_t)__cvta_generic_to_shared(&R), keys_s = (uint32_t)__cvta_generic_to_shared(keys);
    uint32_t queue_s = (uint32_t)__cvta_generic_to_shared(queue), lane_p = lane;
    uint32_t span_s = (uint32_t)__cvta_generic_to_shared(R.span[warp]), span_mg_s = (uint32_t)__cvta_generic_to_shared(R.span_mg[warp]);
    asm volatile("mov.u32 %0, %0;" : "+r"(span_s));
    asm volatile("mov.u32 %0, %0;" : "+r"(span_mg_s));
    asm volatile("mov.u32 %0, %0;" : "+r"(recs_s));
    asm volatile("mov.u32 %0, %0;" : "+r"(keys_s));
    asm volatile("mov.u32 %0, %0;" : "+r"(queue_s));
    asm volatile("mov.u32 %0, %0;" : "+r"(lane_p));
    const uint32_t lt_mask = (1u << lane_p) - 1u, le_mask = 0xffffffffu >> (31u - lane_p);
    auto lds_u32 = [](uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; };
    auto lds_f32 = [](uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; };
    auto sts_u32 = [](uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); };
#define REC(field) (recs_s + (uint32_t)offsetof(TileRecs, field))

#if TR_PHASE_CLOCKS   // experiments: where a CTA's cycles go (warp 0 = a set-up warp, warp 7 = a walker)
    long long ck_setup = 0, ck_scan = 0, ck_walk = 0, ck_tail = 0, ck_job = 0, ck_t = clock64(), ck_rounds = 0, ck_a = 0, ck_b = 0, ck_c = 0, ck_d = 0;
#define CK(acc) { const long long t_ = clock64(); acc += t_ - ck_t; ck_t = t_; }
#else
#define CK(acc)
#endif
    while (true) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(p.tile_ticket, 1u);
        for (uint32_t i = tid; i < TS * TS; i += TILE_THREADS) keys[i] = 0ull;
        if (tid < (TS / 8) * (TS / 8)) R.zmin_blk[tid] = 0.0f;
        __syncthreads();
        if (s_item >= *p.n_jobs) break;
        const uint32_t job = p.tile_order[s_item], item = job & 0x0fffffffu, quadrant = job >> 28;   // 0: the whole tile
        const uint32_t layer = item / p.n_tiles, tile = item - layer * p.n_tiles;
        const uint32_t ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        // the job's pixels: the tile, or one 32x32 quadrant of it (same key array, same stride, a quarter of it in use)
        const int rs = quadrant ? TS / 2 : TS;
        const int tile_x0 = (int)tx * TS + (quadrant ? (int)((quadrant - 1u) & 1u) * (TS / 2) : 0);
        const int tile_y0 = (int)(ty + p.tile_row0) * TS + (quadrant ? (int)((quadrant - 1u) >> 1) * (TS / 2) : 0);
        // the tile's DEPTH_BUCKETS lists are consecutive: one sequence, nearest instances first
        const uint32_t begin = min(p.bin_start[item * DEPTH_BUCKETS], p.bin_capacity), end = min(p.bin_start[(item + 1) * DEPTH_BUCKETS], p.bin_capacity);
        const uint32_t count = end - begin;

        uint32_t e_next = 0;   // this thread's entry of the coming round
        if (tid < ROUND && tid < count) e_next = p.bin_entries[begin + tid];
        CK(ck_job)
        for (uint32_t round = 0; round < count; round += ROUND) {
#if TR_PHASE_CLOCKS
            ck_rounds++;
#endif
            // ---- thread i: set triangle i up, park both forms in shared memory
            uint32_t n_samples = 0, n_box = 0, n_rows = 0;
            if (tid >= TILE_THREADS / 2 && round) {
                // the upper half of the CTA meanwhile refreshes the hierarchical Z (min depth of each 8x8 pixel block) from the
                // depths the previous round left.  The set-up threads may read a block's old or new minimum: depths only ever
                // move nearer, so either is a valid (conservative) bound.  A warp reads 32 consecutive keys of a pixel row with
                // one 64-bit load per lane (conflict-free; the depth words alone sit on 16 banks), eight rows deep, and
                // the minimum of every eight lanes is one block's.
                constexpr uint32_t HALVES = TS / 32, UNITS = (TS / 8) * HALVES;   // a unit: 32 pixels x 8 rows = 4 blocks
                for (uint32_t unit = warp - TILE_THREADS / 64; unit < UNITS; unit += TILE_THREADS / 64) {
                    const uint32_t by = unit / HALVES, half = unit % HALVES;
                    float m = __int_as_float(0x7f800000);
#pragma unroll
                    for (uint32_t k = 0; k < 8; k++) m = fminf(m, __uint_as_float((uint32_t)(keys[(by * 8u + k) * TS + half * 32u + lane] >> 32)));
                    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                    // (an atomic store: the concurrent readers are meant, and compute-sanitizer's racecheck is told so)
                    if ((lane & 7u) == 0u) atomicExch(reinterpret_cast<uint32_t*>(&R.zmin_blk[by * (TS / 8) + half * 4u + (lane >> 3)]), __float_as_uint(m));
                }
                CK(ck_a)
            }
            if (tid < ROUND && round + tid < count) {
                const uint32_t e = e_next;
                if (round + ROUND + tid < count) e_next = p.bin_entries[begin + round + ROUND + tid];
                TriSetup s;
                uint32_t rec_gtid, rec_inst;
                load_trirec(p.trirec + e, s, rec_gtid, rec_inst);
#if TR_PHASE_CLOCKS
                if (s.x_lo + s.y_lo + (int)rec_gtid + (int)__double2loint(s.A[0]) + (int)__float_as_uint(s.W[2]) == -12345) ck_rounds++;
                CK(ck_a)
#endif
                {
                    const int x_lo = max(s.x_lo, tile_x0) - tile_x0, x_hi = min(s.x_hi, tile_x0 + rs - 1) - tile_x0;
                    const int y_lo = max(s.y_lo, tile_y0) - tile_y0, y_hi = min(s.y_hi, tile_y0 + rs - 1) - tile_y0;
                    if (x_lo <= x_hi && y_lo <= y_hi) {
                        const int bw = x_hi - x_lo + 1;
                        n_samples = (uint32_t)(bw * (y_hi - y_lo + 1));
                        R.box[tid] = (uint32_t)x_lo | ((uint32_t)y_lo << 6) | ((uint32_t)(bw - 1) << 12);
                        R.gtid[tid] = e;   // the bin entry is the triangle's work-list id
                        if (CLIP) {
                            const tr_instance* inst = p.instances + rec_inst;
                            const tr_primitive_info* prim = p.prims + __ldg(&inst->primitive_id);
                            const uint32_t rec_slot = find_slot(p, e, p.scalars[0]);   // alpha-clip buckets only
                            R.entry[tid] = make_uint2(rec_slot, e - __ldg(p.work_prefix + rec_slot));
                            R.clip_mat[tid] = (__ldg(&prim->draw_buffer_index) & 1u) ? __ldg(&inst->material_id) : 0xffffffffu;
                        }
                        const double X0 = (double)tile_x0 + 0.5, Y0 = (double)tile_y0 + 0.5;
                        double cl[3], n_a = 0.0, n_b = 0.0, n_c = 0.0, det = 0.0, absdet = 0.0;
                        float max_z = 0.0f, min_w = s.W[0];
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            R.A[i][tid] = s.A[i];
                            R.B[i][tid] = s.B[i];
                            R.C[i][tid] = s.C[i];
                            R.Z[i][tid] = s.Z[i];
                            R.W[i][tid] = s.W[i];
                            cl[i] = s.A[i] * X0 + s.B[i] * Y0 + s.C[i];  // edge value at the tile's first pixel centre
                            R.ea[i][tid] = (float)s.A[i];
                            R.eb[i][tid] = (float)s.B[i];
                            R.ec[i][tid] = (float)cl[i];
                            // |fp32 edge value - double edge value| <= 2^-21 (64(|A|+|B|) + |c|) with a 2x reserve, plus the
                            // rounding of the double evaluation itself (absolute coordinates)
                            const double m = 64.0 * (fabs(s.A[i]) + fabs(s.B[i])) + fabs(cl[i]);
                            const double m_abs = 16384.0 * (fabs(s.A[i]) + fabs(s.B[i])) + fabs(s.C[i]);
                            R.ebound[i][tid] = (float)(m * 4.76837158203125e-7 + m_abs * 1e-15) * 1.0001f;
                            n_a += s.A[i] * (double)s.Z[i];
                            n_b += s.B[i] * (double)s.Z[i];
                            n_c += cl[i] * (double)s.Z[i];
                            det += cl[i] * (double)s.W[i];
                            absdet += fabs(cl[i] * (double)s.W[i]);
                            max_z = fmaxf(max_z, fabsf(s.Z[i]));
                            min_w = fminf(min_w, s.W[i]);
                        }
                        // conservative depth plane: depth(x, y) = sum_i E_i Z_i / sum_i E_i W_i and the denominator is constant
                        float d0 = 0.0f, gx = 0.0f, gy = 0.0f, mg = __int_as_float(0x7f800000);
                        if (min_w > 0.0f && det > 0.0 && absdet < det * 1048576.0) {
                            const double r = 1.0 / det;
                            d0 = (float)(n_c * r);
                            gx = (float)(n_a * r);
                            gy = (float)(n_b * r);
                            mg = (fabsf(d0) + 64.0f * (fabsf(gx) + fabsf(gy))) * 1.9073486e-6f + (max_z / min_w) * 9.536743e-7f;
                            if (!(mg == mg)) mg = __int_as_float(0x7f800000);
                        }
                        R.d0[tid] = d0;
                        R.gx[tid] = gx;
                        R.gy[tid] = gy;
                        R.margin[tid] = mg;
#if TR_PHASE_CLOCKS
                        if (mg == -12345.0f) ck_rounds++;
                        CK(ck_b)
#endif
                        // hierarchical Z: the whole clipped box is behind what the tile already holds
                        const float dmax = d0 + fmaxf(gx * (float)x_lo, gx * (float)x_hi) + fmaxf(gy * (float)y_lo, gy * (float)y_hi) + mg;
                        float zm = __int_as_float(0x7f800000);
                        {   // min over the blocks the clipped box touches: whole block rows at a time (four blocks per load)
                            const int bx0 = x_lo >> 3, bx1 = x_hi >> 3;
                            for (int by = y_lo >> 3; by <= (y_hi >> 3); by++) {
#pragma unroll
                                for (int g = 0; g < TS / 32; g++) {
                                    if (bx1 < 4 * g || bx0 > 4 * g + 3) continue;
                                    const float4 z4 = *reinterpret_cast<const float4*>(&R.zmin_blk[by * (TS / 8) + 4 * g]);
                                    if (bx0 <= 4 * g && 4 * g <= bx1) zm = fminf(zm, z4.x);
                                    if (bx0 <= 4 * g + 1 && 4 * g + 1 <= bx1) zm = fminf(zm, z4.y);
                                    if (bx0 <= 4 * g + 2 && 4 * g + 2 <= bx1) zm = fminf(zm, z4.z);
                                    if (bx0 <= 4 * g + 3 && 4 * g + 3 <= bx1) zm = fminf(zm, z4.w);
                                }
                            }
                        }
                        n_box = n_samples;
#if TR_PHASE_CLOCKS
                        if (zm == -12345.0f) ck_rounds++;
                        CK(ck_c)
#endif
                        if (dmax < zm) n_samples = 0;
                        else n_rows = (uint32_t)(y_hi - y_lo + 1);
                    }
                }
            }
            {   // diagnostics
                const uint32_t a = __reduce_add_sync(0xffffffffu, n_box), b = __reduce_add_sync(0xffffffffu, n_samples);
                if (lane == 0 && a) {
                    atomicAdd(p.stats, (unsigned long long)a);
                    atomicAdd(p.stats + 1, (unsigned long long)b);
                }
            }
            // ---- CTA-wide inclusive scan of the box heights: the unit of the walk is one pixel row of one clipped box
            uint32_t incl = n_rows;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += o;
            }
            if (lane == 31) R.warp_tot[warp] = incl;
#if TR_PREFETCH
            // the coming round's records are pulled towards the SM while this round is walked: the set-up is 64 threads behind a
            // 128-byte gather, and the other six warps of the CTA wait for it
            if (tid < ROUND && round + ROUND + tid < count) {
#if TR_PREFETCH == 1
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p.trirec + e_next));
#else
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.trirec + e_next));
#endif
            }
#endif
            __syncthreads();
            CK(ck_setup)
            uint32_t warp_off = 0, total_rows = 0;
#pragma unroll
            for (uint32_t k = 0; k < TILE_THREADS / 32; k++) {
                const uint32_t t = R.warp_tot[k];
                if (k < warp) warp_off += t;
                total_rows += t;
            }
            if (tid < ROUND) R.off[tid + 1] = warp_off + incl;
            if (tid == 0) {
                R.off[0] = 0;
                R.row_ticket = 0;
            }
            __syncthreads();
            CK(ck_scan)

            // ---- the walk.  Each warp takes 32 box rows at a time.  Step 1, one lane per row: the span of the row that passes
            // the three fp32 edge tests.  fmaf(a, x, v) is monotone in x, so each test holds on a half line and the survivors
            // of a row are ONE interval; its ends come from a division and are then corrected with the very predicate the
            // pixels would have been tested with — the span is exactly the set a per-pixel test would keep (half of the box
            // on average, which the per-pixel walk of round 1 paid for in full).  Step 2, one lane per span pixel (the spans of
            // the 32 rows laid end to end): conservative depth plane against the tile's current depth, survivors queued.
            uint32_t qn = 0, n_exact = 0, n_span = 0;
            while (true) {   // 32 rows at a time, handed out by a ticket: a warp held up by exact evaluations takes fewer
                uint32_t ticket = 0;
                if (lane_p == 0) ticket = atomicAdd(&R.row_ticket, 1u);
                ticket = __shfl_sync(0xffffffffu, ticket, 0);
                // 32 rows per ticket, 16 over the last TR_TAIL_ROWS rows of the round: the round ends when its slowest warp does
                const uint32_t full = total_rows > TR_TAIL_ROWS ? (total_rows - TR_TAIL_ROWS + 31u) >> 5 : 0u;
                const uint32_t rbase = ticket < full ? ticket * 32u : full * 32u + (ticket - full) * 16u;
                const uint32_t rend = min(rbase + (ticket < full ? 32u : 16u), total_rows);
                if (rbase >= total_rows) break;
                const uint32_t r = rbase + lane_p;
                uint32_t len = 0, q0 = 0;
                float rowbase = 0.0f, r_gx = 0.0f, r_mg = 0.0f;
                if (r < rend) {
                    uint32_t lo = 0, hi = ROUND;   // largest j with off[j] <= r
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (lds_u32(REC(off) + 4u * mid) <= r) lo = mid; else hi = mid;
                    }
                    const uint32_t j = lo, rj = recs_s + 4u * j;  // field [i][j] sits at offsetof(field) + 4 (ROUND i + j)
                    const uint32_t box = lds_u32(rj + (uint32_t)offsetof(TileRecs, box));
                    const uint32_t ly = ((box >> 6) & 63u) + (r - lds_u32(rj + (uint32_t)offsetof(TileRecs, off)));
                    int x0 = (int)(box & 63u), x1 = x0 + (int)(box >> 12);   // inclusive
                    const float fy = (float)ly;
#pragma unroll
                    for (uint32_t i = 0; i < 3; i++) {
                        const float a = lds_f32(rj + (uint32_t)offsetof(TileRecs, ea) + 4u * ROUND * i);
                        const float v = fmaf(lds_f32(rj + (uint32_t)offsetof(TileRecs, eb) + 4u * ROUND * i), fy,
                                             lds_f32(rj + (uint32_t)offsetof(TileRecs, ec) + 4u * ROUND * i));
                        const float nt = -lds_f32(rj + (uint32_t)offsetof(TileRecs, ebound) + 4u * ROUND * i);
                        auto pass = [&](int x) { return !(fmaf(a, (float)x, v) < nt); };   // the per-pixel edge test
                        if (a > 0.0f) {          // holds for x >= (-t - v) / a
                            int cand = (int)fminf(fmaxf(ceilf(__fdividef(nt - v, a)), (float)x0), (float)(x1 + 1));
                            while (cand > x0 && pass(cand - 1)) cand--;
                            while (cand <= x1 && !pass(cand)) cand++;
                            x0 = cand;
                        } else if (a < 0.0f) {   // holds for x <= (-t - v) / a
                            int cand = (int)fmaxf(fminf(floorf(__fdividef(nt - v, a)), (float)x1), (float)(x0 - 1));
                            while (cand < x1 && pass(cand + 1)) cand++;
                            while (cand >= x0 && !pass(cand)) cand--;
                            x1 = cand;
                        } else if (a == 0.0f && v < nt) {
                            x1 = x0 - 1;
                        }
                    }
                    if (x1 >= x0) {
                        len = (uint32_t)(x1 - x0 + 1);
                        q0 = j | ((uint32_t)x0 << 8) | (ly << 14);
                        rowbase = fmaf(lds_f32(rj + (uint32_t)offsetof(TileRecs, gy)), fy, lds_f32(rj + (uint32_t)offsetof(TileRecs, d0)));
                        r_gx = lds_f32(rj + (uint32_t)offsetof(TileRecs, gx));
                        r_mg = lds_f32(rj + (uint32_t)offsetof(TileRecs, margin));
                    }
                }
                uint32_t sincl = len;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xffffffffu, sincl, d);
                    if (lane >= (uint32_t)d) sincl += o;
                }
                const uint32_t excl = sincl - len, total = __shfl_sync(0xffffffffu, sincl, 31);
                const uint32_t nz = __ballot_sync(0xffffffffu, len != 0u);
                if (len) {   // the non-empty spans, compacted, in this warp's descriptor table
                    const uint32_t at = span_s + 16u * (uint32_t)__popc(nz & lt_mask);
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(at), "r"(excl), "r"(q0), "r"(__float_as_uint(rowbase)),
                                 "r"(__float_as_uint(r_gx)) : "memory");
                    sts_u32(span_mg_s + 4u * (uint32_t)__popc(nz & lt_mask), __float_as_uint(r_mg));
                }
                __syncwarp();
                n_span += total;
                uint32_t before = 0;   // spans that start before the current window of 32 samples
                for (uint32_t sb = 0; sb < total; sb += 32u) {
                    // bit k: a span starts at sample sb + k.  The span of sample sb + lane = before + starts at or below it - 1.
                    const uint32_t starts = __reduce_or_sync(0xffffffffu, (len != 0u && excl - sb < 32u) ? 1u << (excl - sb) : 0u);
                    const uint32_t sidx = sb + lane_p;
                    bool survive = false;
                    uint32_t q = 0;
                    if (sidx < total) {
                        const uint32_t sp = before + (uint32_t)__popc(starts & le_mask) - 1u;
                        uint32_t d_excl, d_q0, d_base, d_gx;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d_excl), "=r"(d_q0), "=r"(d_base), "=r"(d_gx) : "r"(span_s + 16u * sp));
                        const float mg = lds_f32(span_mg_s + 4u * sp);
                        q = d_q0 + ((sidx - d_excl) << 8);
                        const uint32_t lx = (q >> 8) & 63u, ly = q >> 14;
                        const float cur = __uint_as_float(lds_u32(keys_s + (ly * TS + lx) * 8u + 4u));  // depth half of the key
                        const float dz = fmaf(__uint_as_float(d_gx), (float)lx, __uint_as_float(d_base));
                        survive = !(dz + mg < cur);
                    }
                    before += (uint32_t)__popc(starts);
                    const uint32_t m = __ballot_sync(0xffffffffu, survive);
                    if (survive) sts_u32(queue_s + 4u * (qn + __popc(m & lt_mask)), q);
                    qn += __popc(m);
                    n_exact += __popc(m);
                    __syncwarp();
                    if (qn >= 32u) {
                        const uint32_t mine = lds_u32(queue_s + 4u * lane_p);
                        const uint32_t spill = lane_p + 32u < qn ? lds_u32(queue_s + 4u * (lane_p + 32u)) : 0u;
                        __syncwarp();
                        sts_u32(queue_s + 4u * lane_p, spill);
                        qn -= 32u;
                        exact_sample<TS, CLIP>(p, R, keys, mine, tile_x0, tile_y0);
                        __syncwarp();
                    }
                }
                __syncwarp();   // the descriptor table is rewritten by the next 32 rows
            }
            if (lane_p < qn) exact_sample<TS, CLIP>(p, R, keys, lds_u32(queue_s + 4u * lane_p), tile_x0, tile_y0);
            if (lane == 0 && n_exact) atomicAdd(p.stats + 2, (unsigned long long)n_exact);
            if (lane == 0 && n_span) atomicAdd(p.stats + 3, (unsigned long long)n_span);
            CK(ck_walk)
            __syncthreads();  // the records are rewritten by the next round
            CK(ck_tail)
        }

        __syncthreads();
        unsigned long long* out = p.vis[layer];
        for (uint32_t i = tid; i < TS * TS; i += TILE_THREADS) {
            const int lx = (int)(i & (TS - 1)), ly = (int)(i / TS), px = tile_x0 + lx, py = tile_y0 + ly;
            if (lx < rs && ly < rs && px < (int)p.width && py >= (int)p.y0 && py < (int)p.y1) out[(size_t)py * p.width + px] = keys[i];
        }
    }
#if TR_PHASE_CLOCKS
    CK(ck_job)
    if ((blockIdx.x == 0 || blockIdx.x == 300) && (tid == 0 || tid == 224))
        printf("cta %u warp %u: rounds %lld setup %lld scan %lld walk %lld tail %lld job-overhead %lld | load/refresh %lld math %lld hiz %lld\n", blockIdx.x, warp, ck_rounds, ck_setup, ck_scan,
               ck_walk, ck_tail, ck_job, ck_a, ck_b, ck_c);
#endif
}

// ---- pass C: resolve.  One CTA per 32x8 pixel block, both layers.  The block's distinct winning triangles (a few
// dozen: neighbouring pixels share triangles) are collected in a shared-memory hash set and set up ONCE each, by the first
// threads of the CTA; every pixel then only evaluates its triangle's plane and interpolates.  Setting the triangle up per
// pixel, as the first version did, cost 2.6x the instructions.  Same arithmetic (setup_triangle / eval_pixel), same bits.
constexpr int RES_W = 32, RES_H = 8, RES_TABLE = 1024, RES_CAP = 160;
struct ResolveRec {
    double A[3], B[3], C[3];
    float Z[3], W[3];
    float n[3][3];   // rotation * normal per vertex (lib.rs:356), set-up order
    float uv[3][2];
    uint32_t material_id;
    float scale;
};

__device__ __forceinline__ void resolve_setup(const VisParams& p, uint32_t gtid, ResolveRec& r) {
    TriSetup s;
    uint32_t rec_gtid, inst_id;
    load_trirec(p.trirec + gtid, s, rec_gtid, inst_id);   // pass A1's set-up of this triangle
    const tr_instance* inst = p.instances + inst_id;
    const float4 rot = __ldg(reinterpret_cast<const float4*>(inst) + 1);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.A[k] = s.A[k];
        r.B[k] = s.B[k];
        r.C[k] = s.C[k];
        r.Z[k] = s.Z[k];
        r.W[k] = s.W[k];
        const float* mn = p.normals + (size_t)s.vid[k] * 3;
        const f3 nk = xquat_mul3(rot.x, rot.y, rot.z, rot.w, mk3(__ldg(mn), __ldg(mn + 1), __ldg(mn + 2)));  // lib.rs:356
        r.n[k][0] = nk.x;
        r.n[k][1] = nk.y;
        r.n[k][2] = nk.z;
        r.uv[k][0] = __ldg(p.uvs + (size_t)s.vid[k] * 2);
        r.uv[k][1] = __ldg(p.uvs + (size_t)s.vid[k] * 2 + 1);
    }
    r.material_id = __ldg(&inst->material_id);
    r.scale = __ldg(&inst->transform.translation_and_scale.w);
}

template <bool DERIV>
__device__ __forceinline__ void resolve_write(const VisParams& p, const ResolveRec& r, int layer, uint32_t i, int px, int py) {
    TriSetup s;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        s.A[k] = r.A[k];
        s.B[k] = r.B[k];
        s.C[k] = r.C[k];
        s.Z[k] = r.Z[k];
        s.W[k] = r.W[k];
    }
    float l[3] = {0.f, 0.f, 0.f}, d = 0.0f;
    eval_plane(s, px, py, l, d);  // eval_pixel's arithmetic without its coverage tests: by construction the winning triangle covers this pixel
    p.depth[layer][i] = d;
    p.normal[layer][(size_t)i * 3 + 0] = xadd(xadd(xmul(l[0], r.n[0][0]), xmul(l[1], r.n[1][0])), xmul(l[2], r.n[2][0]));
    p.normal[layer][(size_t)i * 3 + 1] = xadd(xadd(xmul(l[0], r.n[0][1]), xmul(l[1], r.n[1][1])), xmul(l[2], r.n[2][1]));
    p.normal[layer][(size_t)i * 3 + 2] = xadd(xadd(xmul(l[0], r.n[0][2]), xmul(l[1], r.n[1][2])), xmul(l[2], r.n[2][2]));
    const float uv_u = xadd(xadd(xmul(l[0], r.uv[0][0]), xmul(l[1], r.uv[1][0])), xmul(l[2], r.uv[2][0]));
    const float uv_v = xadd(xadd(xmul(l[0], r.uv[0][1]), xmul(l[1], r.uv[1][1])), xmul(l[2], r.uv[2][1]));
    p.uv[layer][(size_t)i * 2 + 0] = uv_u;
    p.uv[layer][(size_t)i * 2 + 1] = uv_v;
    if (DERIV) {  // forward differences to (x+1, y) and (x, y+1) on this triangle's plane
        float dq[6];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            float ln[3], dn;
            dq[k * 2] = dq[k * 2 + 1] = dq[4 + k] = 0.0f;
            if (eval_plane(s, px + (k == 0), py + (k == 1), ln, dn)) {
                const float un = xadd(xadd(xmul(ln[0], r.uv[0][0]), xmul(ln[1], r.uv[1][0])), xmul(ln[2], r.uv[2][0]));
                const float vn = xadd(xadd(xmul(ln[0], r.uv[0][1]), xmul(ln[1], r.uv[1][1])), xmul(ln[2], r.uv[2][1]));
                dq[k * 2] = xsub(un, uv_u);
                dq[k * 2 + 1] = xsub(vn, uv_v);
                dq[4 + k] = xsub(dn, d);
            }
        }
        p.duv[layer][i] = make_float4(dq[0], dq[1], dq[2], dq[3]);
        p.ddepth[layer][i] = make_float2(dq[4], dq[5]);
    }
    p.material_id[layer][i] = r.material_id;
    if (layer == 1) p.scale1[i] = r.scale;
}

#ifndef TR_RESOLVE_CTAS
#define TR_RESOLVE_CTAS 5
#endif
#ifndef TR_RESOLVE_PX
#define TR_RESOLVE_PX 2
#endif
// The pass is a chain of three dependent trips to memory per block (keys -> records -> vertex attributes) and little else, so
// its rate is (pixels in flight per SM) / (the chain's latency): with one pixel per thread 1536 pixels were in flight per SM
// and nothing done to the arithmetic, the set-up or the stores moved its 0.19 ms.  RES_PX pixels per thread (rows RES_H apart
// in a 32 x (RES_H RES_PX) block) put more pixels behind every trip.
constexpr int RES_PX = TR_RESOLVE_PX;
template <bool DERIV>
__global__ void __launch_bounds__(RES_W * RES_H, TR_RESOLVE_CTAS) resolve_kernel(const __grid_constant__ VisParams p) {
    __shared__ uint32_t s_key[RES_TABLE];     // triangle id, 0xffffffff = empty
    __shared__ uint16_t s_idx[RES_TABLE];     // compact index of that triangle
    __shared__ uint32_t s_list[2 * RES_W * RES_H * RES_PX];
    __shared__ uint32_t s_count;
    __shared__ ResolveRec s_rec[RES_CAP];
    const uint32_t tid = threadIdx.x;
    const int px = (int)(blockIdx.x * RES_W + (tid & (RES_W - 1)));
    const int py0 = (int)(p.y0 + blockIdx.y * (RES_H * RES_PX) + tid / RES_W);

    for (uint32_t k = tid; k < RES_TABLE; k += RES_W * RES_H) s_key[k] = 0xffffffffu;
    if (tid == 0) s_count = 0;
    __syncthreads();

    // phase 1: the winning triangle of each of this thread's pixels in each layer goes into the hash set
    uint32_t gtid[RES_PX][2], slot[RES_PX][2];
    bool inside[RES_PX];
    unsigned long long k0[RES_PX], k1[RES_PX];
#pragma unroll
    for (int q = 0; q < RES_PX; q++) {   // all the key loads first: one trip to memory for all of them
        const int py = py0 + q * RES_H;
        inside[q] = px < (int)p.width && py < (int)p.y1;
        const uint32_t i = (uint32_t)py * p.width + (uint32_t)px;
        k0[q] = inside[q] ? p.vis[0][i] : 0ull;
        k1[q] = inside[q] ? p.vis[1][i] : 0ull;
    }
    const uint32_t lane = tid & 31u;
#pragma unroll
    for (int q = 0; q < RES_PX; q++) {
        // depth GREATER against the opaque depth of the shared depth buffer, applied to the nearest transmissive
        // fragment (if that one is hidden, every other one is too)
        if (!(__uint_as_float((uint32_t)(k1[q] >> 32)) > __uint_as_float((uint32_t)(k0[q] >> 32)))) k1[q] = 0ull;
        gtid[q][0] = k0[q] ? 0xffffffffu - (uint32_t)(k0[q] & 0xffffffffull) : 0xffffffffu;
        gtid[q][1] = k1[q] ? 0xffffffffu - (uint32_t)(k1[q] & 0xffffffffull) : 0xffffffffu;
#pragma unroll
        for (int layer = 0; layer < 2; layer++) {
            // a warp is one row of 32 pixels and usually sees one to three triangles: one lane per distinct triangle goes to
            // the table (otherwise all lanes would fight over the same shared-memory word), the rest get its slot by shuffle
            const uint32_t peers = __match_any_sync(0xffffffffu, gtid[q][layer]);
            const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
            uint32_t h = 0;
            if (lane == leader && gtid[q][layer] != 0xffffffffu) {
                h = (gtid[q][layer] * 2654435761u) >> 22;  // 10 bits
                while (true) {
                    const uint32_t old = atomicCAS(&s_key[h], 0xffffffffu, gtid[q][layer]);
                    if (old == 0xffffffffu) {  // first to see this triangle: give it a compact index
                        const uint32_t idx = atomicAdd(&s_count, 1u);
                        s_idx[h] = (uint16_t)idx;
                        s_list[idx] = gtid[q][layer];
                        break;
                    }
                    if (old == gtid[q][layer]) break;
                    h = (h + 1) & (RES_TABLE - 1);
                }
            }
            slot[q][layer] = __shfl_sync(0xffffffffu, h, (int)leader);
        }
    }
    __syncthreads();

    // phase 2: one thread per distinct triangle sets it up (packed into the first warps)
    const uint32_t n_tri = s_count;
    for (uint32_t t = tid; t < min(n_tri, (uint32_t)RES_CAP); t += RES_W * RES_H) resolve_setup(p, s_list[t], s_rec[t]);
    __syncthreads();

    // phase 3: per pixel
#pragma unroll
    for (int q = 0; q < RES_PX; q++) {
        if (!inside[q]) continue;
        const int py = py0 + q * RES_H;
        const uint32_t i = (uint32_t)py * p.width + (uint32_t)px;
#pragma unroll
        for (int layer = 0; layer < 2; layer++) {
            if (gtid[q][layer] == 0xffffffffu) {
                p.depth[layer][i] = 0.0f;
                p.normal[layer][(size_t)i * 3] = 0.0f; p.normal[layer][(size_t)i * 3 + 1] = 0.0f; p.normal[layer][(size_t)i * 3 + 2] = 0.0f;
                p.uv[layer][(size_t)i * 2] = 0.0f; p.uv[layer][(size_t)i * 2 + 1] = 0.0f;
                p.material_id[layer][i] = 0xffffffffu;
                if (layer == 1) p.scale1[i] = 0.0f;
                if (DERIV) {
                    p.duv[layer][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    p.ddepth[layer][i] = make_float2(0.f, 0.f);
                }
                continue;
            }
            const uint32_t idx = s_idx[slot[q][layer]];
            if (idx < (uint32_t)RES_CAP) {
                resolve_write<DERIV>(p, s_rec[idx], layer, i, px, py);
            } else {  // more distinct triangles in this block than records: set this one up privately
                ResolveRec r;
                resolve_setup(p, gtid[q][layer], r);
                resolve_write<DERIV>(p, r, layer, i, px, py);
            }
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
        TR_TRY(c->vis[l].ensure(npx * 8));
        TR_TRY(ensure_layer(c, l, false));
    }
    TR_TRY(ensure_tri_bound(c, "tr_visibility"));
    VisParams p{};
    // tile edge: 64 pixels, or 32 when the band is so small that 64-pixel tiles would leave SMs idle / unbalanced
    {
        const uint32_t t64 = ((c->width + 63) / 64) * ((c->band_y1 - 1) / 64 - c->band_y0 / 64 + 1);
        // measured on 1/4 and 1/8 bands of a 4K frame: 32-pixel tiles balance better but cost more set-up work per
        // triangle than they save, so they are only used when 64-pixel tiles cannot even occupy half the CTA slots
        p.ts = 2 * t64 * 2 < (uint32_t)c->sm_count * TR_TILE_CTAS ? 32 : 64;
        if (const char* e = getenv("TR_TILE_EDGE")) p.ts = atoi(e) == 32 ? 32 : 64;   // experiments
    }
    p.tiles_x = (c->width + p.ts - 1) / p.ts;
    p.tile_row0 = c->band_y0 / p.ts;
    p.tiles_y = (c->band_y1 - 1) / p.ts - p.tile_row0 + 1;
    p.n_tiles = p.tiles_x * p.tiles_y;
    if (p.tiles_x > 256 || p.tiles_y > 256) return fail(TR_ERR_UNSUPPORTED, "tr_visibility: frame or band larger than 256 tiles on a side");
    // (triangle, tile) pairs: a fixed floor, more for big scenes; an overflow is still detected (status bit 1) and reported
    // by every read-back entry point, never silently accepted
    p.bin_capacity = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(BIN_CAPACITY, 6 * c->max_triangles + 64ull * p.n_tiles), 1ull << 27);
    p.rec_capacity = (uint32_t)(c->max_triangles ? c->max_triangles : 1);
    TR_TRY(c->bin_entries.ensure((size_t)p.bin_capacity * sizeof(uint32_t)));
    // per triangle of the work list: room for the 128-byte set-up (at its work-list id) and the 16-byte binning record
    TR_TRY(c->tri_records.ensure((size_t)p.rec_capacity * (sizeof(uint4) + sizeof(TriRec))));
    // state block: [bin_count L][rec_count, ticket, pad, pad][bin_start L+1][bin_cursor L]; the first two parts are zeroed per frame
    const size_t n_lists = (size_t)2 * p.n_tiles * DEPTH_BUCKETS;
    p.n_lists = (uint32_t)n_lists;
    const size_t zero_bytes = (n_lists + 4 + 32) * 4;  // bin_count, rec_count/ticket, scan totals
    TR_TRY(c->bin_state.ensure(zero_bytes + (2 * n_lists + 4) * 4 + (size_t)8 * p.n_tiles * 4));
    if (!c->dev_status.p) {
        TR_TRY(c->dev_status.ensure(64));
        TR_CUDA(cudaMemsetAsync(c->dev_status.p, 0, 64, c->stream));
    }
    TR_CUDA(cudaMemsetAsync(c->bin_state.p, 0, zero_bytes, c->stream));
    uint32_t* st = c->bin_state.as<uint32_t>();
    p.bin_count = st;
    p.rec_count = st + n_lists;
    p.tile_ticket = st + n_lists + 1;
    p.n_jobs = st + n_lists + 2;
    p.scan_totals = st + n_lists + 4;
    p.bin_start = st + n_lists + 4 + 32;
    p.bin_cursor = p.bin_start + n_lists + 4;  // keeps 16-byte alignment (n_lists is a multiple of 16)
    p.tile_order = p.bin_cursor + n_lists;
    p.bin_entries = c->bin_entries.as<uint32_t>();
    p.trirec = c->tri_records.as<TriRec>();
    p.records = reinterpret_cast<uint4*>(p.trirec + p.rec_capacity);
    p.status = c->dev_status.as<uint32_t>();
    p.stats = reinterpret_cast<unsigned long long*>(c->dev_status.as<unsigned char>() + 16);

    p.positions = c->mesh_pos.as<float>();
    p.normals = c->mesh_nrm.as<float>();
    p.uvs = c->mesh_uv.as<float>();
    p.indices = c->mesh_idx.as<uint32_t>();
    p.instances = c->instances.as<tr_instance>();
    p.prims = c->primitives.as<tr_primitive_info>();
    p.visible_ids = c->visible_ids.as<uint32_t>();
    p.work_prefix = c->work_prefix.as<uint32_t>();
    p.scalars = c->d_cull_scalars;
    p.slot_z = c->slot_z.as<float>();
    p.slot_prim = c->slot_first.as<uint4>();
    // (a mesh uploaded between tr_cull and this call leaves K1's chunk bases pointing into the old table: no chunk test then)
    p.chunk_spheres = c->chunk_cull && c->chunks_valid ? c->chunk_spheres.as<float4>() : nullptr;
    {   // clip-space half spaces x >= -w, x <= w and the band's rows (two pixels of slack) as world-space planes, in double
        const float* m = reinterpret_cast<const float*>(&pc.proj_view);   // column-major: element (row r, col k) = m[k * 4 + r]
        auto row = [&](int r, int k) { return (double)m[k * 4 + r]; };
        const double y_lo = 2.0 * ((double)c->band_y0 - 2.0) / (double)c->height - 1.0, y_hi = 2.0 * ((double)c->band_y1 + 2.0) / (double)c->height - 1.0;
        for (int q = 0; q < 4; q++) {
            double pl[4];
            for (int k = 0; k < 4; k++)
                pl[k] = q == 0 ? row(3, k) + row(0, k) : q == 1 ? row(3, k) - row(0, k) : q == 2 ? row(1, k) - y_lo * row(3, k) : y_hi * row(3, k) - row(1, k);
            p.cull_plane[q] = make_float4((float)pl[0], (float)pl[1], (float)pl[2], (float)pl[3]);
            p.cull_plane_norm[q] = (float)(sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]) * 1.0001);
        }
    }
    memcpy(&p.proj_view, &pc.proj_view, sizeof(mat4));
    {
        const float* m = reinterpret_cast<const float*>(&pc.proj_view);  // column-major: element (row r, col k) = m[k * 4 + r]
        p.row_y_norm = sqrtf(m[1] * m[1] + m[5] * m[5] + m[9] * m[9]) * 1.0001f;
        p.row_w_norm = sqrtf(m[3] * m[3] + m[7] * m[7] + m[11] * m[11]) * 1.0001f;
        p.band_cull = (c->band_y0 != 0 || c->band_y1 != c->height) ? 1u : 0u;
    }
    p.width = c->width;
    p.height = c->height;
    p.y0 = c->band_y0;
    p.y1 = c->band_y1;
    for (int l = 0; l < 2; l++) {
        p.vis[l] = c->vis[l].as<unsigned long long>();
        p.depth[l] = c->layer[l].depth.as<float>();
        p.normal[l] = c->layer[l].normal.as<float>();
        p.uv[l] = c->layer[l].uv.as<float>();
        p.material_id[l] = c->layer[l].material_id.as<uint32_t>();
    }
    p.scale1 = c->layer[1].scale.as<float>();
    p.materials = nullptr;
    if (c->prims_alpha_clip) {
        if (!c->n_materials) return fail(TR_ERR_STATE, "tr_visibility: alpha-clip primitives need the materials (tr_set_materials)");
        TR_TRY(upload_texture_table(c));
        p.materials = c->materials.as<tr_material_info>();
        p.textures = c->tex_table.as<TexDesc>();
        p.n_textures = c->n_textures;
    }
    for (int l = 0; l < 2; l++) {
        p.duv[l] = c->materials_textured ? c->layer[l].duv.as<float4>() : nullptr;
        p.ddepth[l] = c->materials_textured ? c->layer[l].ddepth.as<float2>() : nullptr;
    }

    const size_t tile_smem = (size_t)p.ts * p.ts * 8 + sizeof(TileRecs);
    auto tile_kernel = p.materials ? (p.ts == 64 ? raster_tiles_kernel<64, true> : raster_tiles_kernel<32, true>)
                                   : (p.ts == 64 ? raster_tiles_kernel<64, false> : raster_tiles_kernel<32, false>);
    TR_CUDA(cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
    int per_sm = 0;
    TR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_kernel, TILE_THREADS, tile_smem));
    if (per_sm < 1) return fail(TR_ERR_CUDA, "raster_tiles_kernel does not fit on an SM (smem %zu)", tile_smem);
    uint32_t tile_grid = (uint32_t)(c->sm_count * per_sm);
    // fewer than three jobs per CTA slot: the pass would last as long as its heaviest tile — let heavy tiles split
    p.split = p.ts == 64 && 2 * p.n_tiles < 3 * tile_grid ? 1u : 0u;
    if (const char* e = getenv("TR_TILE_SPLIT")) p.split = atoi(e) && p.ts == 64 ? 1u : 0u;   // experiments
    if (tile_grid > 8 * p.n_tiles) tile_grid = 8 * p.n_tiles;

    p.list_prefix = p.work_prefix;
    p.list_slots = nullptr;
    p.list_block_entry = c->block_entry.as<uint32_t>();
    p.band_block_entry = c->block_entry.as<uint32_t>() + c->max_triangles / 256 + 2;
    p.list_scalars = p.scalars;
    const bool no_table = getenv("TR_NO_BLOCK_ENTRY") != nullptr;   // A/B switch: search the prefix instead
    int extra_launch = 0;
    if (p.band_cull && c->n_instances <= (1u << 17)) {  // the single-CTA filter is meant for visible sets of this size
        TR_TRY(c->band_list.ensure(((size_t)c->n_instances * 2 + 1 + 4) * 4));
        p.band_slots = c->band_list.as<uint32_t>();
        p.band_prefix = p.band_slots + c->n_instances;
        p.band_scalars = p.band_prefix + c->n_instances + 1;
        band_filter_kernel<<<1, 1024, 0, c->stream>>>(p);
        p.list_prefix = p.band_prefix;
        p.list_slots = p.band_slots;
        p.list_block_entry = p.band_block_entry;
        p.list_scalars = p.band_scalars;
        extra_launch = 1;
    }
    if (no_table) p.list_block_entry = nullptr;
    bin_count_kernel<<<c->sm_count * TR_BIN_WAVES * TR_BIN_CTAS, 256, 0, c->stream>>>(p);
    // one CTA per 8192 lists, at most SCAN_CTAS: a small band's few thousand lists are not worth a chain of 32 waiting CTAs
    bin_scan_kernel<<<std::max(1u, std::min<uint32_t>(SCAN_CTAS, (p.n_lists + 8191u) / 8192u)), 1024, 0, c->stream>>>(p);
    bin_fill_kernel<<<c->sm_count * 2 * TR_BIN_CTAS, 256, 0, c->stream>>>(p);
    tile_kernel<<<tile_grid, TILE_THREADS, tile_smem, c->stream>>>(p);
    const dim3 res_grid((c->width + RES_W - 1) / RES_W, (c->band_y1 - c->band_y0 + RES_H * RES_PX - 1) / (RES_H * RES_PX), 1);
    if (c->materials_textured) resolve_kernel<true><<<res_grid, RES_W * RES_H, 0, c->stream>>>(p);
    else resolve_kernel<false><<<res_grid, RES_W * RES_H, 0, c->stream>>>(p);
    count_launches(5 + extra_launch);
    TR_CUDA(cudaGetLastError());
    for (int l = 0; l < 2; l++) {
        c->layer[l].valid = true;
        c->layer[l].has_position = false;
    }
    return TR_OK;
}

}  // namespace tr

// k_cull.cu — K1: frustum culling + ordered compaction + draw demultiplexing (sm_100a).
//
// Reference: `frustum_culling` + `cull` shader/src/lib.rs:411-469 (64-thread groups, one atomic
// increment per visible instance, main.rs:1762) and `demultiplex_draws` lib.rs:471-517
// (atomic slot per non-empty primitive, main.rs:1837).  The reference only COUNTS instances and
// appends draws in atomic-race order; here one launch classifies every instance with the exact
// arithmetic of `cull` (bit-identical visibility bits), counts per primitive, and emits the
// visible-instance ids in ascending order with a warp-ballot + shuffle block scan chained across
// CTAs by a decoupled look-back (single pass, no host round trip).  The same scan carries the
// running triangle count, which is the work list of the visibility kernel (K3).  A second tiny
// launch writes the four indirect draw lists in ascending primitive order.
#include "tr_internal.h"

using namespace trd;

namespace {

constexpr int CULL_THREADS = 256;
constexpr uint64_t ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_MASK = 3ull << 62;
constexpr int VIS_BITS = 24;  // descriptor = status(2) | triangles(38) | visible(24)

struct CullParams {
    const tr_instance* inst;
    uint32_t n_inst;
    const tr_primitive_info* prims;
    tr_culling_push_constants pc;
    uint32_t* ticket;
    unsigned long long* desc;   // [n_blocks]
    uint32_t* instance_counts;  // [n_prims]
    uint32_t* visible_ids;      // [n_inst]
    uint32_t* work_prefix;      // [n_inst + 1] exclusive triangle prefix per visible slot
    uint32_t* scalars;          // [0] n_visible, [1] total triangles (low 32), [2..5] draw_counts, [6] ~min / [7] max bits of slot_z
    float* slot_z;              // [n_inst] nearest view-space depth of each visible instance (front-to-back binning in K3)
    uint4* slot_prim;           // [n_inst] (first_index, draw_buffer_index, first culling chunk, 0) of the slot's primitive: K3 skips the
                                // instance -> primitive hop
    uint32_t* block_entry;      // [work-list triangles / 256 + 1] the slot that holds triangle 256 b: K3's binning pass starts there instead of
                                // searching the prefix (twelve dependent loads per warp)
    const uint32_t* prim_chunk_base;   // [n_prims + 1] first culling chunk of each primitive (ensure_chunks)
};

// shader/src/lib.rs:442-469, exact regime
__device__ __forceinline__ bool cull(float4 sphere, float4 ts, float4 rot, const tr_culling_push_constants& pc, float& nearest_z) {
    f3 center = mk3(sphere.x, sphere.y, sphere.z);
    // Similarity * Vec3, shared-structs lib.rs:235-241
    center = xadd3(mk3(ts.x, ts.y, ts.z), xscale3(xquat_mul3(rot.x, rot.y, rot.z, rot.w, center), ts.w));
    const mat4& view = *reinterpret_cast<const mat4*>(&pc.view);
    f4 c = xmat4_mul(view, center.x, center.y, center.z, 1.0f);
    const float cz = -c.z;  // lib.rs:452
    const float radius = xmul(sphere.w, ts.w);
    nearest_z = fmaxf(cz - radius, pc.z_near);
    bool visible = xadd(cz, radius) > pc.z_near;
    visible &= xsub(xmul(cz, pc.frustum_x_xz.y), xmul(fabsf(c.x), pc.frustum_x_xz.x)) < radius;
    visible &= xsub(xmul(cz, pc.frustum_y_yz.y), xmul(fabsf(c.y), pc.frustum_y_yz.x)) < radius;
    return !visible;
}

__global__ void __launch_bounds__(CULL_THREADS) cull_kernel(const __grid_constant__ CullParams p) {
    __shared__ uint32_t s_bid;
    __shared__ unsigned long long s_warp[CULL_THREADS / 32];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_bid = atomicAdd(p.ticket, 1u);  // blocks are numbered in arrival order: predecessors always run
    __syncthreads();
    const uint32_t bid = s_bid;
    const uint32_t i = bid * CULL_THREADS + tid;

    bool visible = false;
    uint32_t tris = 0, first_index = 0, draw_buffer = 0, chunk_base = 0;
    float nearest_z = 0.0f;
    if (i < p.n_inst) {
        const float4* q = reinterpret_cast<const float4*>(p.inst + i);
        const float4 ts = __ldg(q), rot = __ldg(q + 1);
        const uint4 ids = __ldg(reinterpret_cast<const uint4*>(q + 2));
        const uint32_t prim = ids.x;
        const float4* pq = reinterpret_cast<const float4*>(p.prims + prim);
        const float4 sphere = __ldg(pq);
        const uint4 pinfo = __ldg(reinterpret_cast<const uint4*>(pq + 1));
        visible = !cull(sphere, ts, rot, p.pc, nearest_z);
        if (visible) {
            tris = pinfo.y / 3u;
            first_index = pinfo.z;
            draw_buffer = pinfo.x;
            chunk_base = __ldg(p.prim_chunk_base + prim);
            atomicAdd(p.instance_counts + prim, 1u);  // lib.rs:437-439
        }
    }

    // block-exclusive scan of (triangles << 24 | visible)
    unsigned long long v = ((unsigned long long)tris << VIS_BITS) | (visible ? 1ull : 0ull);
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < CULL_THREADS / 32; w++) {
        const unsigned long long t = s_warp[w];
        if (w < warp) warp_off += t;
        block_total += t;
    }
    const unsigned long long local_excl = warp_off + incl - v;

    // decoupled look-back across CTAs (warp 0)
    if (warp == 0) {
        if (lane == 0) {
            const unsigned long long d = (bid == 0 ? ST_PREFIX : ST_AGG) | block_total;
            atomicExch(p.desc + bid, d);
        }
        unsigned long long excl = 0;
        if (bid > 0) {
            int j = (int)bid - 1;
            while (true) {
                const int idx = j - lane;
                unsigned long long d;
                do {
                    d = idx >= 0 ? *reinterpret_cast<volatile unsigned long long*>(p.desc + idx) : ST_PREFIX;
                } while (__any_sync(0xffffffffu, (d & ST_MASK) == 0));
                const uint32_t pmask = __ballot_sync(0xffffffffu, (d & ST_MASK) == ST_PREFIX);
                const int first = pmask ? __ffs(pmask) - 1 : 32;
                unsigned long long val = lane <= first ? (d & ~ST_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                excl += val;
                if (first < 32) break;
                j -= 32;
            }
            if (lane == 0) atomicExch(p.desc + bid, ST_PREFIX | (excl + block_total));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    const unsigned long long base = s_excl + local_excl;
    if (visible) {
        const uint32_t slot = (uint32_t)(base & ((1ull << VIS_BITS) - 1));
        p.visible_ids[slot] = i;
        p.work_prefix[slot] = (uint32_t)(base >> VIS_BITS);
        p.slot_z[slot] = nearest_z;
        p.slot_prim[slot] = make_uint4(first_index, draw_buffer, chunk_base, 0u);
        const uint32_t t0 = (uint32_t)(base >> VIS_BITS);
        for (uint32_t b = (t0 + 255u) >> 8; (b << 8) < t0 + tris; b++) p.block_entry[b] = slot;
    }
    // depth range of the visible set (positive floats order like their bit patterns)
    const uint32_t zb = visible ? __float_as_uint(nearest_z) : 0u;
    const uint32_t zmax = __reduce_max_sync(0xffffffffu, zb), zmin_c = __reduce_max_sync(0xffffffffu, visible ? ~zb : 0u);
    if (lane == 0 && zmax) {
        atomicMax(p.scalars + 7, zmax);
        atomicMax(p.scalars + 6, zmin_c);
    }
    if (bid == gridDim.x - 1 && tid == 0) {
        const unsigned long long total = s_excl + block_total;
        const uint32_t nv = (uint32_t)(total & ((1ull << VIS_BITS) - 1));
        p.scalars[0] = nv;
        p.scalars[1] = (uint32_t)(total >> VIS_BITS);
        p.work_prefix[nv] = (uint32_t)(total >> VIS_BITS);
    }
}

struct DemuxParams {
    const tr_primitive_info* prims;
    const uint32_t* instance_counts;
    uint32_t n_prims;
    tr_draw_indexed_indirect_command* draws[4];
    uint32_t* draw_counts;  // [4]
};

// shader/src/lib.rs:471-517, one CTA, ascending primitive order per bucket
__global__ void __launch_bounds__(1024) demux_kernel(const __grid_constant__ DemuxParams p) {
    __shared__ unsigned long long s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t chunk = 0; chunk < p.n_prims; chunk += 1024) {
        const uint32_t d = chunk + tid;
        uint32_t count = 0, bucket = 0;
        tr_primitive_info prim{};
        if (d < p.n_prims) {
            count = p.instance_counts[d];
            prim = p.prims[d];
            bucket = prim.draw_buffer_index < 3u ? prim.draw_buffer_index : 3u;  // match arm `_`, lib.rs:511-516
        }
        const unsigned long long v = count ? 1ull << (16 * bucket) : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long off = 0, total = 0;
        for (int w = 0; w < 32; w++) {
            if (w < warp) off += s_warp[w];
            total += s_warp[w];
        }
        if (count) {
            // chunk-local rank (16 bits per bucket) + the running per-bucket total kept in draw_counts
            const uint32_t local = (uint32_t)(((off + incl - v) >> (16 * bucket)) & 0xffffu);
            tr_draw_indexed_indirect_command c;
            c.index_count = prim.index_count;
            c.instance_count = count;
            c.first_index = prim.first_index;
            c.vertex_offset = 0;
            c.first_instance = prim.first_instance;
            p.draws[bucket][p.draw_counts[bucket] + local] = c;
        }
        __syncthreads();
        if (tid < 4) p.draw_counts[tid] += (uint32_t)((total >> (16 * tid)) & 0xffffu);
        __syncthreads();
    }
}

}  // namespace

namespace tr {

int32_t launch_cull(tr_ctx* c, const tr_culling_push_constants& pc) {
    if (!c->n_instances || !c->n_primitives) return fail(TR_ERR_STATE, "tr_cull: instances and primitives must be set");
    TR_TRY(validate_scene(c, "tr_cull"));  // the kernel indexes prims[] and instance_counts[] by instance.primitive_id
    const uint32_t n_blocks = (c->n_instances + CULL_THREADS - 1) / CULL_THREADS;
    // state block: [ticket | pad][desc x n_blocks][instance_counts x n_prims][scalars x 8] -> one memset per frame
    const size_t desc_off = 16, counts_off = desc_off + (size_t)n_blocks * 8;
    const size_t scalars_off = counts_off + (((size_t)c->n_primitives * 4 + 15) & ~(size_t)15);
    const size_t state_bytes = scalars_off + 32;
    TR_TRY(c->cull_scalars.ensure(state_bytes));
    TR_TRY(c->visible_ids.ensure((size_t)c->n_instances * 4));
    TR_TRY(c->work_prefix.ensure(((size_t)c->n_instances + 1) * 4));
    TR_TRY(ensure_tri_bound(c, "tr_cull"));
    // one table for K1's list, one behind it for a band's own list (band_filter_kernel)
    TR_TRY(c->block_entry.ensure((size_t)(c->max_triangles / 256 + 2) * 2 * 4));
    TR_TRY(c->slot_z.ensure((size_t)c->n_instances * 4));
    TR_TRY(c->slot_first.ensure((size_t)c->n_instances * 16));
    TR_TRY(ensure_chunks(c));
    for (int b = 0; b < 4; b++) TR_TRY(c->draws[b].ensure((size_t)c->n_primitives * sizeof(tr_draw_indexed_indirect_command)));
    unsigned char* st = c->cull_scalars.as<unsigned char>();
    // zeroing the instance count / draw count buffers, main.rs:1669-1700
    TR_CUDA(cudaMemsetAsync(st, 0, state_bytes, c->stream));
    c->d_instance_counts = reinterpret_cast<uint32_t*>(st + counts_off);  // views into the state block
    c->d_cull_scalars = reinterpret_cast<uint32_t*>(st + scalars_off);

    CullParams p;
    p.inst = c->instances.as<tr_instance>();
    p.n_inst = c->n_instances;
    p.prims = c->primitives.as<tr_primitive_info>();
    p.pc = pc;
    p.ticket = reinterpret_cast<uint32_t*>(st);
    p.desc = reinterpret_cast<unsigned long long*>(st + desc_off);
    p.instance_counts = reinterpret_cast<uint32_t*>(st + counts_off);
    p.visible_ids = c->visible_ids.as<uint32_t>();
    p.work_prefix = c->work_prefix.as<uint32_t>();
    p.block_entry = c->block_entry.as<uint32_t>();
    p.slot_z = c->slot_z.as<float>();
    p.slot_prim = c->slot_first.as<uint4>();
    p.prim_chunk_base = c->prim_chunk_base.as<uint32_t>();
    p.scalars = reinterpret_cast<uint32_t*>(st + scalars_off);
    cull_kernel<<<n_blocks, CULL_THREADS, 0, c->stream>>>(p);
    count_launches(2);
    TR_CUDA(cudaGetLastError());

    DemuxParams d;
    d.prims = p.prims;
    d.instance_counts = p.instance_counts;
    d.n_prims = c->n_primitives;
    for (int b = 0; b < 4; b++) d.draws[b] = c->draws[b].as<tr_draw_indexed_indirect_command>();
    d.draw_counts = p.scalars + 2;
    demux_kernel<<<1, 1024, 0, c->stream>>>(d);
    TR_CUDA(cudaGetLastError());
    c->cull_valid = true;
    return TR_OK;
}

}  // namespace tr

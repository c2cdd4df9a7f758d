// k_mips.cu — K5: the opaque-frame mip chain in ONE launch (sm_100a).
//
// Reference: `generate_mips(opaque_sampled_hdr_framebuffer, w, h, levels)`
// src/main.rs:2054-2063, level count src/main.rs:2590-2592; the body is a
// vkCmdBlitImage(LINEAR) chain in ash-opinionated-abstractions (not in tree),
// i.e. one full read of every level from DRAM per level.  Filter rule and
// fp16 rounding points: SURVEY.md Appendix E / oracle/mips.c — bit-exact.
//
// B200 design: every CTA reduces one 64x64 tile of level 0 through all the
// levels whose source has even width AND height (2x2 box, the tile never
// needs a neighbour): level 0 is read from HBM exactly once with 128-bit
// loads, level 1/2 are produced in registers, deeper levels in shared memory.
// The few remaining small levels (first odd-sized source onwards, <= 240x135
// at 4K: 10.7 k texels in seven levels, each depending on the one before) are
// built by a second launch of ONE thread-block cluster: eight CTAs x 1024
// threads, a hardware cluster barrier between levels, out of L2.  (Round 1 let
// the last CTA of the first launch do it alone: 256 threads, ~13 of the pass's
// 40 us, and a ticket + fence in every CTA.)
#include <cooperative_groups.h>

#include "tr_internal.h"

namespace cg = cooperative_groups;
using namespace trd;

namespace {

struct MipParams {
    uint2* base;
    uint32_t levels;
    uint32_t n_local;  // levels 1..n_local are tile-local
    uint32_t w[tr::kMaxLevels], h[tr::kMaxLevels], off[tr::kMaxLevels];
    uint32_t* counter;
};

struct h4 { float x, y, z, w; };

__device__ __forceinline__ h4 unpack(uint2 v) {
    f4 t = unpack_rgba16f(v);
    h4 r; r.x = t.x; r.y = t.y; r.z = t.z; r.w = t.w;
    return r;
}
// lerp in the exact regime: a + (b - a) * t   (oracle/mips.c)
__device__ __forceinline__ float xlerp(float a, float b, float t) { return xadd(a, xmul(xsub(b, a), t)); }
__device__ __forceinline__ uint2 filter4(uint2 a00, uint2 a10, uint2 a01, uint2 a11, float fx, float fy) {
    h4 t00 = unpack(a00), t10 = unpack(a10), t01 = unpack(a01), t11 = unpack(a11);
    float r = xlerp(xlerp(t00.x, t10.x, fx), xlerp(t01.x, t11.x, fx), fy);
    float g = xlerp(xlerp(t00.y, t10.y, fx), xlerp(t01.y, t11.y, fx), fy);
    float b = xlerp(xlerp(t00.z, t10.z, fx), xlerp(t01.z, t11.z, fx), fy);
    float a = xlerp(xlerp(t00.w, t10.w, fx), xlerp(t01.w, t11.w, fx), fy);
    return pack_rgba16f(r, g, b, a);
}
__device__ __forceinline__ uint2 box(uint2 a00, uint2 a10, uint2 a01, uint2 a11) { return filter4(a00, a10, a01, a11, 0.5f, 0.5f); }

__device__ __forceinline__ void axis_setup(uint32_t d, uint32_t src, uint32_t dst, uint32_t& i0, uint32_t& i1, float& frac) {
    float scale = xdiv((float)src, (float)dst);
    float p = xsub(xmul(xadd((float)d, 0.5f), scale), 0.5f);
    float fl = floorf(p);
    int i = (int)fl, hi = (int)src - 1;
    i0 = (uint32_t)min(max(i, 0), hi);
    i1 = (uint32_t)min(max(i + 1, 0), hi);
    frac = xsub(p, fl);
}

__global__ void __launch_bounds__(256) mip_kernel(const __grid_constant__ MipParams p) {
    __shared__ uint2 s2[16][17];
    __shared__ uint2 s3[8][9];
    __shared__ uint2 s4[4][5];
    __shared__ uint2 s5[2][3];
    const int tid = threadIdx.x;

    if (p.n_local >= 1) {
        const uint32_t w0 = p.w[0], h0 = p.h[0], w1 = p.w[1], h1 = p.h[1];
        const uint2* l0 = p.base + p.off[0];
        uint2* l1 = p.base + p.off[1];
        const uint32_t tx = tid & 15, ty = tid >> 4;
        const uint32_t x = blockIdx.x * 64 + tx * 4, y = blockIdx.y * 64 + ty * 4;
        // 4x4 block of level 0 -> 2x2 of level 1 (128-bit loads, two texels each)
        uint4 r[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int hx = 0; hx < 2; hx++) {
                const uint32_t xx = x + hx * 2, yy = y + j;
                r[j][hx] = (xx < w0 && yy < h0) ? __ldcs(reinterpret_cast<const uint4*>(l0 + (size_t)yy * w0 + xx))
                                                : make_uint4(0, 0, 0, 0);
            }
        uint2 q[2][2];
#pragma unroll
        for (int qy = 0; qy < 2; qy++)
#pragma unroll
            for (int qx = 0; qx < 2; qx++) {
                const uint4 top = r[qy * 2][qx], bot = r[qy * 2 + 1][qx];
                q[qy][qx] = box(make_uint2(top.x, top.y), make_uint2(top.z, top.w), make_uint2(bot.x, bot.y),
                                make_uint2(bot.z, bot.w));
            }
        const uint32_t x1 = x >> 1, y1 = y >> 1;
#pragma unroll
        for (int qy = 0; qy < 2; qy++) {
            const uint32_t yy = y1 + qy;
            if (yy < h1) {
                if ((w1 & 1u) == 0u && x1 + 1 < w1) {
                    *reinterpret_cast<uint4*>(l1 + (size_t)yy * w1 + x1) =
                        make_uint4(q[qy][0].x, q[qy][0].y, q[qy][1].x, q[qy][1].y);
                } else {
                    if (x1 < w1) l1[(size_t)yy * w1 + x1] = q[qy][0];
                    if (x1 + 1 < w1) l1[(size_t)yy * w1 + x1 + 1] = q[qy][1];
                }
            }
        }
        if (p.n_local >= 2) {
            const uint32_t w2 = p.w[2], h2 = p.h[2];
            const uint32_t x2 = x >> 2, y2 = y >> 2;
            const uint2 v2 = box(q[0][0], q[0][1], q[1][0], q[1][1]);
            if (x2 < w2 && y2 < h2) p.base[p.off[2] + (size_t)y2 * w2 + x2] = v2;
            s2[ty][tx] = v2;
        }
        if (p.n_local >= 3) {
            __syncthreads();
            const uint32_t ox = tid & 7, oy = tid >> 3;
            if (tid < 64) {
                const uint32_t gx = blockIdx.x * 8 + ox, gy = blockIdx.y * 8 + oy;
                const uint2 v = box(s2[oy * 2][ox * 2], s2[oy * 2][ox * 2 + 1], s2[oy * 2 + 1][ox * 2], s2[oy * 2 + 1][ox * 2 + 1]);
                if (gx < p.w[3] && gy < p.h[3]) p.base[p.off[3] + (size_t)gy * p.w[3] + gx] = v;
                s3[oy][ox] = v;
            }
        }
        if (p.n_local >= 4) {
            __syncthreads();
            const uint32_t ox = tid & 3, oy = tid >> 2;
            if (tid < 16) {
                const uint32_t gx = blockIdx.x * 4 + ox, gy = blockIdx.y * 4 + oy;
                const uint2 v = box(s3[oy * 2][ox * 2], s3[oy * 2][ox * 2 + 1], s3[oy * 2 + 1][ox * 2], s3[oy * 2 + 1][ox * 2 + 1]);
                if (gx < p.w[4] && gy < p.h[4]) p.base[p.off[4] + (size_t)gy * p.w[4] + gx] = v;
                s4[oy][ox] = v;
            }
        }
        if (p.n_local >= 5) {
            __syncthreads();
            const uint32_t ox = tid & 1, oy = tid >> 1;
            if (tid < 4) {
                const uint32_t gx = blockIdx.x * 2 + ox, gy = blockIdx.y * 2 + oy;
                const uint2 v = box(s4[oy * 2][ox * 2], s4[oy * 2][ox * 2 + 1], s4[oy * 2 + 1][ox * 2], s4[oy * 2 + 1][ox * 2 + 1]);
                if (gx < p.w[5] && gy < p.h[5]) p.base[p.off[5] + (size_t)gy * p.w[5] + gx] = v;
                s5[oy][ox] = v;
            }
        }
        if (p.n_local >= 6) {
            __syncthreads();
            if (tid == 0) {
                const uint2 v = box(s5[0][0], s5[0][1], s5[1][0], s5[1][1]);
                if (blockIdx.x < p.w[6] && blockIdx.y < p.h[6]) p.base[p.off[6] + (size_t)blockIdx.y * p.w[6] + blockIdx.x] = v;
            }
        }
    }

}

constexpr int TAIL_CLUSTER = 8, TAIL_THREADS = 1024;
__global__ void __cluster_dims__(TAIL_CLUSTER, 1, 1) __launch_bounds__(TAIL_THREADS) mip_tail_kernel(const __grid_constant__ MipParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t t = cluster.block_rank() * TAIL_THREADS + threadIdx.x, nt = TAIL_CLUSTER * TAIL_THREADS;
    for (uint32_t l = p.n_local + 1; l < p.levels; l++) {
        const uint32_t sw = p.w[l - 1], sh = p.h[l - 1], dw = p.w[l], dh = p.h[l];
        const uint2* src = p.base + p.off[l - 1];
        uint2* dst = p.base + p.off[l];
        for (uint32_t i = t; i < dw * dh; i += nt) {
            const uint32_t dy = i / dw, dx = i - dy * dw;
            uint32_t x0, x1, y0, y1;
            float fx, fy;
            axis_setup(dx, sw, dw, x0, x1, fx);
            axis_setup(dy, sh, dh, y0, y1, fy);
            // .cg: the source level was written by other CTAs (the first launch, or this cluster one barrier ago)
            dst[i] = filter4(__ldcg(src + (size_t)y0 * sw + x0), __ldcg(src + (size_t)y0 * sw + x1), __ldcg(src + (size_t)y1 * sw + x0),
                             __ldcg(src + (size_t)y1 * sw + x1), fx, fy);
        }
        __threadfence();
        cluster.sync();   // barrier.cluster, release / acquire: the level is complete and visible to all eight CTAs
    }
}

}  // namespace

namespace tr {

int32_t launch_generate_mips(uint2* pyramid, uint32_t levels, const uint32_t* w, const uint32_t* h, const uint32_t* off,
                             uint32_t* counter, int sm_count, cudaStream_t s) {
    (void)sm_count;
    if (levels <= 1) return TR_OK;
    MipParams p;
    p.base = pyramid;
    p.levels = levels;
    p.counter = counter;
    for (uint32_t i = 0; i < kMaxLevels; i++) {
        p.w[i] = i < levels ? w[i] : 0;
        p.h[i] = i < levels ? h[i] : 0;
        p.off[i] = i < levels ? off[i] : 0;
    }
    uint32_t n_local = 0;
    while (n_local + 1 < levels && n_local < 6 && (w[n_local] & 1u) == 0 && (h[n_local] & 1u) == 0) n_local++;
    p.n_local = n_local;
    dim3 grid(1, 1, 1);
    if (n_local >= 1) grid = dim3((w[0] + 63) / 64, (h[0] + 63) / 64, 1);
    int launches = 0;
    if (n_local >= 1) {
        mip_kernel<<<grid, 256, 0, s>>>(p);
        launches++;
    }
    if (n_local + 1 < levels) {
        mip_tail_kernel<<<TAIL_CLUSTER, TAIL_THREADS, 0, s>>>(p);
        launches++;
    }
    count_launches(launches);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

}  // namespace tr

// k_shade.cu — K4 opaque shading and K6 transmissive shading (sm_100a).
//
// Reference behaviour:
//   K4  `fragment`              shader/src/lib.rs:164-249 + lighting.rs:145-220
//   K6  `fragment_transmission` shader/src/lib.rs:37-162  + lighting.rs:13-95
// B200 design (DESIGN.md "K4/K6"):
//   * persistent CTAs; every WARP walks its own strided sequence of 32-pixel runs of the band in flattened
//     row-major order, so every G-buffer plane of a run is ONE contiguous range in HBM;
//   * each plane of the warp's next runs is staged into shared memory by the TMA engine (cp.async.bulk +
//     mbarrier complete_tx, SASS UBLKCP) through the warp's own 4-deep ring, so HBM latency is hidden without
//     needing occupancy and no warp ever waits for another one;
//   * lights are converted once per CTA into a compact shared-memory table;
//     the per-pixel cluster lists (ascending light id) are walked with a
//     warp-wide sorted merge (redux.min) so a warp stays converged while every
//     pixel still sums its own cluster's lights in ascending id order — the
//     same order as the oracle, which makes the result independent of tiling;
//   * results are packed to RGBA16F and written with 128-bit stores (two
//     pixels per store, lanes paired by shuffle): even lanes write the
//     hdr_framebuffer, odd lanes write the sampled opaque mip 0 (the reference
//     writes the same value to both targets, lib.rs:247-248).  With the
//     peer-store path the odd lanes write that band into every peer GPU's
//     mip 0 over NVLink, which is the all-gather fused into the shading kernel.
//   * ray-queried shadows (lighting.rs:22-32, 64-71, 154-165, 186-195): this file is compiled a second time as
//     k_shade_shadow.cu with TR_SHADE_SHADOW=1.  That translation unit holds the shadow pass (one thread per pixel
//     traces the sun ray and one ray per light of the pixel's cluster list, tr_device_accel.cuh, and stores the
//     occluded-ray bits) and the shading kernels instantiated to read those bits; the instantiations without ray
//     queries are untouched by it.
#include <type_traits>

#include "tr_internal.h"

#ifndef TR_SHADE_SHADOW
#define TR_SHADE_SHADOW 0
#endif

using namespace trd;

namespace {

constexpr int TILE = 256;   // threads per CTA
constexpr int CHUNK = 32;   // pixels per work item: one warp, one pixel per lane
constexpr int WARPS = TILE / 32;
constexpr int STAGES = 4;
constexpr int MAX_SMEM_LIGHTS = 768;
constexpr int kHeaderBytes = 512 + WARPS * TR_MAX_LIGHTS_PER_CLUSTER * 4;   // barriers + claims, then one light list per warp

struct LightS {  // 64 B, shared-memory form of shared_structs::Light: a point light is the first 40 bytes
    float px, py, pz;
    uint32_t is_spot;
    float er, er2, eg, eg2;        // every colour channel twice: the packed accumulators take (c, c) pairs (tr_device_pbr.cuh)
    float eb, eb2, sx, sy;         // spotlight: direction, cos(outer angle), 1 / epsilon — only read for spotlights
    float sz, cos_outer, inv_eps, pad;
};
static_assert(sizeof(LightS) == 64, "LightS layout");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <bool TRANS, bool HAS_POS>
struct StageLayout {
    static constexpr int kDepth = 0;
    static constexpr int kNormal = kDepth + CHUNK * 4;
    static constexpr int kMat = kNormal + CHUNK * 12;
    static constexpr int kScale = kMat + CHUNK * 4;
    static constexpr int kPos = kScale + (TRANS ? CHUNK * 4 : 0);
    static constexpr int kBytes = kPos + (HAS_POS ? CHUNK * 12 : 0);
};

__device__ __forceinline__ LightS make_light_s(const tr_light* lights, uint32_t i) {
    const float4* q = reinterpret_cast<const float4*>(lights + i);
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    LightS l;
    l.pad = 0.0f;
    l.px = a.x; l.py = a.y; l.pz = a.z;
    l.er = l.er2 = b.x; l.eg = l.eg2 = b.y; l.eb = l.eb2 = b.z;
    l.sx = c.x; l.sy = c.y; l.sz = c.z;
    l.is_spot = c.w != 0.0f;                       // Light::is_a_spotlight, shared-structs lib.rs:125-127
    l.cos_outer = l.is_spot ? (float)cos((double)c.w) : 0.0f;
    l.inv_eps = l.is_spot ? 1.0f / a.w : 0.0f;     // spotlight_factor divides by epsilon, lib.rs:137
    return l;
}

// shader/src/lib.rs:88-98 / 205-215 in the exact regime (this is a discrete decision)
__device__ __forceinline__ uint32_t cluster_index(float fx, float fy, float depth, const tr_uniforms& u) {
    uint32_t cx = f32_as_u32(xdiv(fx, u.cluster_size_in_pixels.x));
    uint32_t cy = f32_as_u32(xdiv(fy, u.cluster_size_in_pixels.y));
    const tr_light_cluster_coefficients& c = u.light_clustering_coefficients;
    // linear_depth, shared-structs lib.rs:54-58
    float depth_range = xsub(xmul(2.0f, xsub(1.0f, depth)), 1.0f);
    float lin = xdiv(xmul(xmul(2.0f, c.z_near), c.z_far),
                     xsub(xadd(c.z_far, c.z_near), xmul(depth_range, xsub(c.z_far, c.z_near))));
    // get_depth_slice, lib.rs:61-63
    float v = xadd(xmul(xlog2_spec(lin), c.scale), c.bias);
    uint32_t cz = f32_as_u32(fmaxf(v, 0.0f));
    return cz * u.num_clusters.x * u.num_clusters.y + cy * u.num_clusters.x + cx;
}

#ifndef TR_LIGHT_UNROLL
#define TR_LIGHT_UNROLL 1
#endif
constexpr int kLightUnroll = TR_LIGHT_UNROLL;
#ifndef TR_SHADE_CTAS_OPAQUE
#define TR_SHADE_CTAS_OPAQUE 3
#endif
#ifndef TR_SHADE_CTAS_TRANS
#define TR_SHADE_CTAS_TRANS 3
#endif
template <bool TRANS, bool HAS_POS, bool F32OUT, bool TEX, bool SHADOW>
__global__ void __launch_bounds__(TILE, TEX ? 2 : (TRANS ? TR_SHADE_CTAS_TRANS : TR_SHADE_CTAS_OPAQUE)) shade_kernel(const __grid_constant__ tr::ShadeLaunch p) {
    using L = StageLayout<TRANS, HAS_POS>;
    extern __shared__ __align__(128) unsigned char smem[];
    // Every warp runs its own pipeline: a ring of STAGES small stages (32 pixels of every G-buffer plane each), its own
    // `full` mbarriers, its own sequence of 32-pixel runs of the band (strided over all warps of the grid).  Lane 0 refills
    // a stage as soon as its warp has read it.  Nothing in the loop synchronises warps with each other, so a warp whose
    // pixels see long light lists never holds the others up (with one 256-pixel tile per CTA and a barrier per tile, ~1 of
    // the ~5 resident warps per scheduler sat at that barrier).
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem) + warp * STAGES;   // this warp's STAGES barriers (bytes 0-255 for all warps)
    uint32_t* s_refs = reinterpret_cast<uint32_t*>(smem + 512) + warp * TR_MAX_LIGHTS_PER_CLUSTER;   // this warp's light list of the run
    unsigned char* stage_base = smem + kHeaderBytes + (size_t)warp * STAGES * L::kBytes;
    LightS* s_lights = reinterpret_cast<LightS*>(smem + kHeaderBytes + (size_t)WARPS * STAGES * L::kBytes);

    const uint32_t n_px = p.px_end - p.px_begin;
    const uint32_t n_tiles = (n_px + CHUNK - 1) / CHUNK;
    // runs are handed out dynamically (one atomic per run on p.chunk_counter): a fixed strided assignment leaves the
    // slowest of the ~3500 resident warps ~15 % behind the average at the end of the kernel (73 runs per warp, cost
    // proportional to the light count), and the kernel ends with the slowest warp
    uint32_t* s_claim = reinterpret_cast<uint32_t*>(smem + 256) + warp * STAGES;   // the run each stage of this warp holds (bytes 256-383)
    const bool lights_in_smem = p.n_lights <= MAX_SMEM_LIGHTS;
    uint32_t lights_saddr = smem_u32(s_lights);
    asm volatile("mov.u32 %0, %0;" : "+r"(lights_saddr));  // opaque to the optimiser: keep it in a register instead of re-deriving it per light

    if (lane == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (lights_in_smem)
        for (uint32_t i = tid; i < p.n_lights; i += TILE) s_lights[i] = make_light_s(p.lights, i);
    __syncthreads();

    auto tile_start = [&](uint32_t t) { return p.px_begin + t * CHUNK; };
    auto tile_count = [&](uint32_t t) { return min((uint32_t)CHUNK, p.px_end - tile_start(t)); };
    auto tile_bulk = [&](uint32_t t) { return ((tile_start(t) & 3u) == 0u) && ((tile_count(t) & 3u) == 0u); };

    auto issue = [&](uint32_t t, int s) {  // lane 0 only
        if (t >= n_tiles || !tile_bulk(t)) return;
        const uint32_t start = tile_start(t), n = tile_count(t);
        unsigned char* sb = stage_base + s * L::kBytes;
        uint32_t bytes = n * (4 + 12 + 4) + (TRANS ? n * 4 : 0) + (HAS_POS ? n * 12 : 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(sb + L::kDepth, p.depth + start, n * 4, &full[s]);
        bulk_g2s(sb + L::kNormal, p.normal + (size_t)start * 3, n * 12, &full[s]);
        bulk_g2s(sb + L::kMat, p.material_id + start, n * 4, &full[s]);
        if (TRANS) bulk_g2s(sb + L::kScale, p.scale + start, n * 4, &full[s]);
        if (HAS_POS) bulk_g2s(sb + L::kPos, p.position + (size_t)start * 3, n * 12, &full[s]);
    };

    if (lane == 0)
        for (int s = 0; s < STAGES; s++) {
            const uint32_t t = atomicAdd(p.chunk_counter, 1u);
            s_claim[s] = t;
            issue(t, s);
        }
    __syncwarp();

    uint32_t phase_bits = 0;
    int stage = 0;
    while (true) {
        const uint32_t t = s_claim[stage];
        if (t >= n_tiles) break;   // claims only grow: nothing valid is left in the other stages either
        const uint32_t start = tile_start(t), n = tile_count(t);
        unsigned char* sb = stage_base + stage * L::kBytes;
        float* s_depth = reinterpret_cast<float*>(sb + L::kDepth);
        float* s_normal = reinterpret_cast<float*>(sb + L::kNormal);
        uint32_t* s_mat = reinterpret_cast<uint32_t*>(sb + L::kMat);
        float* s_scale = reinterpret_cast<float*>(sb + L::kScale);
        float* s_pos = reinterpret_cast<float*>(sb + L::kPos);

        if (tile_bulk(t)) {
            mbar_wait(&full[stage], (phase_bits >> stage) & 1u);
            phase_bits ^= 1u << stage;
        } else {  // ragged run (unaligned start or size): plain loads into the warp's own stage
            if (lane < n) {
                s_depth[lane] = p.depth[start + lane];
                s_mat[lane] = p.material_id[start + lane];
                if (TRANS) s_scale[lane] = p.scale[start + lane];
                for (int k = 0; k < 3; k++) {
                    s_normal[lane * 3 + k] = p.normal[(size_t)(start + lane) * 3 + k];
                    if (HAS_POS) s_pos[lane * 3 + k] = p.position[(size_t)(start + lane) * 3 + k];
                }
            }
            __syncwarp();
        }

        // ------------------------------------------------------------ per-pixel prologue
        const bool active = lane < n;
        const uint32_t g = start + lane;
        const float depth = active ? s_depth[lane] : 0.0f;
        const bool covered = active && depth != 0.0f;

        PixelShading ps;   // TEX: kept for the epilogue; otherwise dead after the prologue (the epilogue re-derives the colour factors)
        LoopPixel lp;
        LoopSums sums;
        sums.clear();
        f3 pos = mk3(0.f, 0.f, 0.f), emission = mk3(0.f, 0.f, 0.f);
        const tr_material_info* mat = nullptr;
        uint32_t my_count = 0, my_base = 0;
        float model_scale = 1.0f;
        float roughness_px = 0.0f, transmission_px = 0.0f, thickness_px = 0.0f;  // after their textures (lib.rs:71-77, 120-124)
        uint32_t occl0 = 0, occl1 = 0, occl2 = 0, occl3 = 0;  // SHADOW: occluded-ray bits by position in the cluster's list

        if (covered) {
            const uint32_t py = g / p.width, px = g - py * p.width;
            const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
            if (HAS_POS) {
                pos = mk3(s_pos[lane * 3], s_pos[lane * 3 + 1], s_pos[lane * 3 + 2]);
            } else {
                // G-buffer decode (oracle: decode_position in oracle/shade.c), exact regime
                float ndc_x = xsub(xmul(xdiv(fx, (float)p.width), 2.0f), 1.0f);
                float ndc_y = xsub(xmul(xdiv(fy, (float)p.height), 2.0f), 1.0f);
                f4 h = xmat4_mul(p.inv_proj_view, ndc_x, ndc_y, depth, 1.0f);
                pos = mk3(xdiv(h.x, h.w), xdiv(h.y, h.w), xdiv(h.z, h.w));
            }
            mat = p.materials + s_mat[lane];
            float4 dfac = __ldg(reinterpret_cast<const float4*>(&mat->diffuse_factor));
            const float4 emis = __ldg(reinterpret_cast<const float4*>(&mat->emissive_factor));
            const float4 scol = __ldg(reinterpret_cast<const float4*>(&mat->specular_colour_factor));
            MaterialParams mp;  // get_material_params, lighting.rs:261-301
            mp.metallic = __ldg(&mat->metallic_factor);
            mp.perceptual_roughness = __ldg(&mat->roughness_factor);
            mp.index_of_refraction = __ldg(&mat->index_of_refraction);
            mp.specular_colour = mk3(scol.x, scol.y, scol.z);
            mp.specular_factor = __ldg(&mat->specular_factor);
            emission = mk3(emis.x, emis.y, emis.z);
            if (TRANS) {
                transmission_px = __ldg(&mat->transmission_factor);
                thickness_px = __ldg(&mat->thickness_factor);
            }

            f3 view_pos = mk3(p.view_position[0], p.view_position[1], p.view_position[2]);
            f3 v = xnormalize3(xsub3(view_pos, pos));                                    // lib.rs:196-197
            const f3 n_in = mk3(s_normal[lane * 3], s_normal[lane * 3 + 1], s_normal[lane * 3 + 2]);
            f3 nrm;
            if (TEX) {
                // texture-mapped material (exact regime, oracle/shade.c): lib.rs:66-77,120-124,190-194; lighting.rs:222-313
                const int32_t* tx = &mat->textures.diffuse;
                TextureSampler ts;
                ts.textures = p.textures;
                ts.n_textures = p.n_textures;
                ts.u = __ldg(p.uv + (size_t)g * 2);
                ts.v = __ldg(p.uv + (size_t)g * 2 + 1);
                ts.duv = __ldg(p.duv + g);
                const int32_t t_diffuse = __ldg(tx + 0), t_mr = __ldg(tx + 1), t_normal = __ldg(tx + 2), t_emissive = __ldg(tx + 3);
                const int32_t t_specular = __ldg(tx + 7), t_spec_colour = __ldg(tx + 8);
                if (t_diffuse != -1) {
                    const f4 smp = ts.sample(t_diffuse);
                    dfac.x = xmul(dfac.x, smp.x); dfac.y = xmul(dfac.y, smp.y); dfac.z = xmul(dfac.z, smp.z); dfac.w = xmul(dfac.w, smp.w);
                }
                f3 dpos_dx = mk3(0.f, 0.f, 0.f), dpos_dy = mk3(0.f, 0.f, 0.f);
                if (!HAS_POS && t_normal != -1) {  // world positions of the right / lower neighbour on this pixel's triangle
                    const float2 dd = __ldg(p.ddepth + g);
                    const float nx1 = xsub(xmul(xdiv((float)(px + 1) + 0.5f, (float)p.width), 2.0f), 1.0f);
                    const float ny0 = xsub(xmul(xdiv(fy, (float)p.height), 2.0f), 1.0f);
                    const float nx0 = xsub(xmul(xdiv(fx, (float)p.width), 2.0f), 1.0f);
                    const float ny1 = xsub(xmul(xdiv((float)(py + 1) + 0.5f, (float)p.height), 2.0f), 1.0f);
                    const f4 hx = xmat4_mul(p.inv_proj_view, nx1, ny0, xadd(depth, dd.x), 1.0f);
                    const f4 hy = xmat4_mul(p.inv_proj_view, nx0, ny1, xadd(depth, dd.y), 1.0f);
                    dpos_dx = xsub3(mk3(xdiv(hx.x, hx.w), xdiv(hx.y, hx.w), xdiv(hx.z, hx.w)), pos);
                    dpos_dy = xsub3(mk3(xdiv(hy.x, hy.w), xdiv(hy.y, hy.w), xdiv(hy.z, hy.w)), pos);
                }
                if (TRANS) {
                    const int32_t t_transmission = __ldg(tx + 5), t_thickness = __ldg(tx + 6);
                    if (t_transmission != -1) transmission_px = xmul(transmission_px, ts.sample(t_transmission).x);
                    if (t_thickness != -1) thickness_px = xmul(thickness_px, ts.sample(t_thickness).y);
                }
                nrm = calculate_normal(n_in, t_normal, ts, dpos_dx, dpos_dy);
                if (t_mr != -1) {
                    const f4 smp = ts.sample(t_mr);
                    mp.metallic = xmul(mp.metallic, smp.z);  // "These two are switched!", lighting.rs:272-276
                    mp.perceptual_roughness = xmul(mp.perceptual_roughness, smp.y);
                }
                if (t_spec_colour != -1) {
                    const f4 smp = ts.sample(t_spec_colour);
                    mp.specular_colour = mk3(xmul(mp.specular_colour.x, smp.x), xmul(mp.specular_colour.y, smp.y), xmul(mp.specular_colour.z, smp.z));
                }
                if (t_specular != -1) mp.specular_factor = xmul(mp.specular_factor, ts.sample(t_specular).w);
                if (t_emissive != -1) {
                    const f4 smp = ts.sample(t_emissive);
                    emission = mk3(xmul(emission.x, smp.x), xmul(emission.y, smp.y), xmul(emission.z, smp.z));
                }
            } else {
                nrm = xnormalize3(n_in);  // lighting.rs:229
            }
            mp.diffuse_colour = mk3(dfac.x, dfac.y, dfac.z);
            roughness_px = mp.perceptual_roughness;
            ps = make_pixel_shading(mp, nrm, v, TRANS);
            lp = make_loop_pixel(ps, pos);
            if (TRANS) model_scale = s_scale[lane];

            const uint32_t cluster = cluster_index(fx, fy, depth, p.uniforms);
            if (cluster < p.n_clusters) {
                my_count = __ldg(p.cluster_counts + cluster);
                my_base = cluster * TR_MAX_LIGHTS_PER_CLUSTER;
            }

            // sun, lighting.rs:37-53 / 171-177 (factor == 1.0 without ray queries)
            f3 sun_dir = mk3(p.uniforms.sun_dir.x, p.uniforms.sun_dir.y, p.uniforms.sun_dir.z);
            f3 sun_int = mk3(p.uniforms.sun_intensity.x, p.uniforms.sun_intensity.y, p.uniforms.sun_intensity.z);
            // the sun's direction is given, so its exact chain starts at the halfway vector; same adaptive rule as the clustered lights
            float sun_factor = 1.0f;
            if (SHADOW) {
                occl0 = __ldg(p.shadow_mask + g);
                occl1 = __ldg(p.shadow_mask + (size_t)p.shadow_plane + g);
                occl2 = __ldg(p.shadow_mask + (size_t)p.shadow_plane * 2 + g);
                occl3 = __ldg(p.shadow_mask + (size_t)p.shadow_plane * 3 + g);
                // opaque pass: factor.max(0.1), lighting.rs:154-165; transmissive pass: the factor as it is, :25-35
                if (__ldg(p.shadow_mask + (size_t)p.shadow_plane * 4 + g) & 1u) sun_factor = TRANS ? 0.0f : 0.1f;
            }
            if (!SHADOW || sun_factor != 0.0f) {
                auto exact_sun = [&]() { return sun_dir; };
                light_lean<TRANS>(lp, exact_sun, sun_dir, dot3(lp.n, sun_dir), dup_colour(sun_int), sun_factor, sums);
            }
        }

        // the staged planes are only read above: refill the stage right away with the run STAGES turns ahead
        __syncwarp();
        if (lane == 0) {
            const uint32_t nt = atomicAdd(p.chunk_counter, 1u);
            s_claim[stage] = nt;
            issue(nt, stage);
        }

        // ------------------------------------------------------------ clustered lights
        const uint32_t* const my_list = p.cluster_indices + my_base;
        // the loop is instantiated twice so that the (launch-uniform) choice between the shared-memory light table and
        // the global fallback for > MAX_SMEM_LIGHTS lights costs nothing per light
        auto light_loop = [&](auto in_smem_tag) {
            constexpr bool IN_SMEM = decltype(in_smem_tag)::value;
            // a light record by its shared-memory byte address (IN_SMEM) or its id (global fallback)
            auto shade_light = [&](uint32_t ref) {
                float4 q0;
                Colour2 col;
                if (IN_SMEM) {
                    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(ref));
                    asm("ld.shared.v2.b64 {%0,%1}, [%2+16];" : "=l"(col.r), "=l"(col.g) : "r"(ref));
                    asm("ld.shared.b64 %0, [%1+32];" : "=l"(col.b) : "r"(ref));
                } else {
                    const float4* g4 = reinterpret_cast<const float4*>(p.lights + ref);
                    const float4 a = __ldg(g4), b = __ldg(g4 + 1), c = __ldg(g4 + 2);
                    q0 = make_float4(a.x, a.y, a.z, __uint_as_float(c.w != 0.0f ? 1u : 0u));
                    col = dup_colour(mk3(b.x, b.y, b.z));
                }
                point_light_lean<TRANS>(lp, mk3(q0.x, q0.y, q0.z), col, [&](const f3& dir, float& factor) {
                    if (!TRANS && __float_as_uint(q0.w) != 0u) {  // lighting.rs:201-203; the transmissive loop has no spotlight factor (:58-92)
                        float sx, sy, sz, cos_outer, inv_eps;
                        if (IN_SMEM) {
                            asm("ld.shared.v2.f32 {%0,%1}, [%2+40];" : "=f"(sx), "=f"(sy) : "r"(ref));
                            asm("ld.shared.v2.f32 {%0,%1}, [%2+48];" : "=f"(sz), "=f"(cos_outer) : "r"(ref));
                            asm("ld.shared.f32 %0, [%1+56];" : "=f"(inv_eps) : "r"(ref));
                        } else {
                            const LightS l = make_light_s(p.lights, ref);
                            sx = l.sx; sy = l.sy; sz = l.sz; cos_outer = l.cos_outer; inv_eps = l.inv_eps;
                        }
                        const float theta = -dot3(dir, mk3(sx, sy, sz));
                        factor *= fmaxf((theta - cos_outer) * inv_eps, 0.0f);
                    }
                }, sums);
            };
            auto light_ref = [&](uint32_t id) { return IN_SMEM ? lights_saddr + id * (uint32_t)sizeof(LightS) : id; };
            auto occluded_word = [&](uint32_t i) { return i < 64u ? (i < 32u ? occl0 : occl1) : (i < 96u ? occl2 : occl3); };
            // 93-96 % of the warps of the 4K workload have all their covered pixels in ONE cluster: the list is then walked
            // with warp-uniform indices (no merge, no per-lane cursor); uncovered lanes just compute along.  The lanes fetch
            // 32 list entries at once (one coalesced load) and hand them round by shuffle: one issue slot per light.
            const uint32_t key = covered ? my_base : 0xffffffffu;
            const uint32_t first = __reduce_min_sync(0xffffffffu, key);
            if (__all_sync(0xffffffffu, key == first || key == 0xffffffffu)) {
                if (first != 0xffffffffu) {
                    const uint32_t count = min(__reduce_max_sync(0xffffffffu, covered ? my_count : 0u), (uint32_t)TR_MAX_LIGHTS_PER_CLUSTER);
                    const uint32_t* list = p.cluster_indices + first;
                    // the lanes fetch the list (coalesced) and park it in the warp's shared-memory slot as light-record references
                    for (uint32_t k = lane; k < count; k += 32u) s_refs[k] = light_ref(__ldg(list + k));
                    __syncwarp();
                    // the list is walked through one shared-memory address register (index + base would be re-derived per
                    // light), one entry ahead, so the entry -> light record -> arithmetic chain starts a light early.  One light
                    // per trip: unrolled by two, the transmissive kernel's loop (with its exact-regime patches) no longer fits
                    // the instruction cache (ncu: no_instruction stalls 0.2 -> 1.9 per issue).
                    uint32_t a = smem_u32(s_refs);
                    const uint32_t a_end = a + count * 4u;
                    auto entry = [&]() {   // reads one entry past the list on the last trip: still this CTA's shared memory
                        uint32_t ref;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ref) : "r"(a));
                        return ref;
                    };
                    uint32_t ref = entry();
                    if (SHADOW) {
                        for (uint32_t i = 0; a != a_end; i++) {
                            a += 4u;
                            const uint32_t nxt = entry();
                            if (!((occluded_word(i) >> (i & 31u)) & 1u)) shade_light(ref);  // an occluded light adds nothing
                            ref = nxt;
                        }
                    } else {
#pragma unroll 1
                        while (a != a_end) {
                            a += 4u;
                            const uint32_t nxt = entry();
                            shade_light(ref);
                            ref = nxt;
                        }
                    }
                    __syncwarp();
                }
                return;
            }
            // mixed warp: every lane walks its own cluster's list (ascending light ids) through a private cursor; the warp
            // takes the smallest pending id each turn, so all lanes stay converged and each pixel still sums in ascending id order
            uint32_t my_i = 0;
            const uint32_t my_n = min(my_count, (uint32_t)TR_MAX_LIGHTS_PER_CLUSTER);
            uint32_t next = my_n ? __ldg(my_list) : 0xffffffffu;
            while (true) {
                const uint32_t m = __reduce_min_sync(0xffffffffu, next);
                if (m == 0xffffffffu) break;
                if (next == m) {
                    const bool occluded = SHADOW && ((occluded_word(my_i) >> (my_i & 31u)) & 1u);
                    my_i++;
                    const uint32_t upcoming = my_i < my_n ? __ldg(my_list + my_i) : 0xffffffffu;  // issued early: hidden behind the BRDF
                    if (!occluded) shade_light(light_ref(m));
                    next = upcoming;
                }
            }
        };
        if (lights_in_smem) light_loop(std::true_type{});
        else light_loop(std::false_type{});

        // ------------------------------------------------------------ epilogue
        float4 out = make_float4(0.0f, 0.0f, 0.0f, 1.0f);  // clear colour, main.rs:1592-1602
        if (covered) {
            f3 base = ps.base, c_diff_pi = ps.c_diff_pi, f0 = ps.f0, df = ps.df;
            if (!TEX) {
                // untextured materials: re-derive the colour factors, which are only needed after the light loop, instead of
                // keeping them in registers across it (registers the loop can use)
                const float4 dfac = __ldg(reinterpret_cast<const float4*>(&mat->diffuse_factor));
                const float4 emis = __ldg(reinterpret_cast<const float4*>(&mat->emissive_factor));
                const float4 scol = __ldg(reinterpret_cast<const float4*>(&mat->specular_colour_factor));
                MaterialParams mp;
                mp.diffuse_colour = mk3(dfac.x, dfac.y, dfac.z);
                mp.metallic = __ldg(&mat->metallic_factor);
                mp.perceptual_roughness = 0.0f;
                mp.index_of_refraction = __ldg(&mat->index_of_refraction);
                mp.specular_colour = mk3(scol.x, scol.y, scol.z);
                mp.specular_factor = __ldg(&mat->specular_factor);
                base = mp.diffuse_colour;
                colour_factors(mp, f0, df, c_diff_pi);
                emission = mk3(emis.x, emis.y, emis.z);
            }
            f3 diff, spec, trans;
            finish_sums(sums, f0, df, c_diff_pi, base, lp.a2, lp.at2, TRANS, diff, spec, trans);
            if (TRANS) {
                const float4 acol = __ldg(reinterpret_cast<const float4*>(&mat->attenuation_colour));
                IblVolumeRefractionParams ip;
                ip.material_params.diffuse_colour = base;
                ip.material_params.metallic = 0.0f;
                ip.material_params.perceptual_roughness = roughness_px;
                ip.material_params.index_of_refraction = __ldg(&mat->index_of_refraction);
                ip.material_params.specular_colour = mk3(0.f, 0.f, 0.f);
                ip.material_params.specular_factor = 0.0f;
                ip.normal = lp.n;
                ip.view = lp.v;
                ip.position = pos;
                ip.thickness = thickness_px;                                // lib.rs:120-124
                ip.model_scale = model_scale;
                ip.attenuation_distance = __ldg(&mat->attenuation_distance);
                ip.attenuation_colour = mk3(acol.x, acol.y, acol.z);
                trans = add3(trans, ibl_volume_refraction(ip, p.proj_view, p.log2_size_x, p.pyramid, p.lut, f0, df));
                const float tf = transmission_px;
                f3 real_t = scale3(trans, tf);                              // lib.rs:157
                diff = lerp3(diff, real_t, tf);                             // lib.rs:159
            }
            out.x = diff.x + spec.x + emission.x;                           // lib.rs:161 / 239
            out.y = diff.y + spec.y + emission.y;
            out.z = diff.z + spec.z + emission.z;
        }

        const uint2 packed = pack_rgba16f(out.x, out.y, out.z, out.w);
        if (!TRANS) {
            // 128-bit paired stores: even lanes -> hdr, odd lanes -> sampled opaque target(s)
            const uint32_t other_x = __shfl_xor_sync(0xffffffffu, packed.x, 1);
            const uint32_t other_y = __shfl_xor_sync(0xffffffffu, packed.y, 1);
            const bool vec_ok = (start & 1u) == 0u;
            if (active) {
                if (F32OUT) p.hdr_f32[g] = out;
                const bool has_partner = vec_ok && (((lane & 1u) != 0u) || (lane + 1u < n));
                if (has_partner) {
                    if ((lane & 1u) == 0u) {
                        *reinterpret_cast<uint4*>(p.hdr + g) = make_uint4(packed.x, packed.y, other_x, other_y);
                    } else {
                        const uint4 v = make_uint4(other_x, other_y, packed.x, packed.y);
#pragma unroll 1
                        for (int k = 0; k < p.n_opaque; k++) *reinterpret_cast<uint4*>(p.opaque[k] + g - 1) = v;
                    }
                } else {
                    p.hdr[g] = packed;
#pragma unroll 1
                    for (int k = 0; k < p.n_opaque; k++) p.opaque[k][g] = packed;
                }
            }
        } else {
            if (covered) {  // the transmission render pass LOADs hdr; only covered pixels are written
                if (F32OUT) p.hdr_f32[g] = out;
                p.hdr[g] = packed;
            }
        }

        stage = stage + 1 == STAGES ? 0 : stage + 1;
    }
}

constexpr bool kShadow = TR_SHADE_SHADOW != 0;

template <bool TRANS, bool HAS_POS, bool F32OUT, bool TEX>
int32_t launch_variant(const tr::ShadeLaunch& p, int sm_count, cudaStream_t s) {
    if (kShadow != (p.shadow_mask != nullptr)) return tr::fail(TR_ERR_STATE, "shade kernel variant and shadow mask disagree");
    using L = StageLayout<TRANS, HAS_POS>;
    const uint32_t n_px = p.px_end - p.px_begin;
    if (n_px == 0) return TR_OK;
    const uint32_t n_tiles = (n_px + TILE - 1) / TILE;   // CTAs worth of 32-pixel runs
    const uint32_t n_smem_lights = p.n_lights <= (uint32_t)MAX_SMEM_LIGHTS ? p.n_lights : 0u;
    const size_t smem = kHeaderBytes + (size_t)WARPS * STAGES * L::kBytes + (size_t)n_smem_lights * sizeof(LightS);
    auto kern = shade_kernel<TRANS, HAS_POS, F32OUT, TEX, kShadow>;
    TR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    TR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TILE, smem));
    if (per_sm < 1) return tr::fail(TR_ERR_CUDA, "shade kernel does not fit on an SM (smem %zu)", smem);
    uint32_t grid = (uint32_t)(sm_count * per_sm);
    if (grid > n_tiles) grid = n_tiles;
    TR_CUDA(cudaMemsetAsync(p.chunk_counter, 0, sizeof(uint32_t), s));
    kern<<<grid, TILE, smem, s>>>(p);
    tr::count_launches(1);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}

template <bool TRANS>
int32_t launch_any(const tr::ShadeLaunch& p, int sm_count, cudaStream_t s) {
    const bool pos = p.position != nullptr, f32 = p.hdr_f32 != nullptr;
    if (p.textures != nullptr) {  // some material binds a texture
        if (pos && f32) return launch_variant<TRANS, true, true, true>(p, sm_count, s);
        if (pos) return launch_variant<TRANS, true, false, true>(p, sm_count, s);
        if (f32) return launch_variant<TRANS, false, true, true>(p, sm_count, s);
        return launch_variant<TRANS, false, false, true>(p, sm_count, s);
    }
    if (pos && f32) return launch_variant<TRANS, true, true, false>(p, sm_count, s);
    if (pos) return launch_variant<TRANS, true, false, false>(p, sm_count, s);
    if (f32) return launch_variant<TRANS, false, true, false>(p, sm_count, s);
    return launch_variant<TRANS, false, false, false>(p, sm_count, s);
}

}  // namespace

#if TR_SHADE_SHADOW
// The shadow pass: one thread per pixel of the band.  Ray set-up in the exact regime, in the oracle's operation order
// (orc_shadow_mask_frame): origin = the fragment's world position, direction / t_max from
// light_direction_and_attenuation (glam-pbr lib.rs:12-23), the sun with t_max 10 000.
template <bool HAS_POS>
__global__ void __launch_bounds__(128) shadow_mask_kernel(const __grid_constant__ tr::ShadeLaunch p) {
    const uint32_t g = p.px_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.px_end) return;
    const float depth = __ldg(p.depth + g);
    uint32_t m[5] = {0u, 0u, 0u, 0u, 0u};
    if (depth != 0.0f) {
        const uint32_t py = g / p.width, px = g - py * p.width;
        const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
        f3 pos;
        if (HAS_POS) {
            pos = mk3(__ldg(p.position + (size_t)g * 3), __ldg(p.position + (size_t)g * 3 + 1), __ldg(p.position + (size_t)g * 3 + 2));
        } else {
            float ndc_x = xsub(xmul(xdiv(fx, (float)p.width), 2.0f), 1.0f);
            float ndc_y = xsub(xmul(xdiv(fy, (float)p.height), 2.0f), 1.0f);
            f4 h = xmat4_mul(p.inv_proj_view, ndc_x, ndc_y, depth, 1.0f);
            pos = mk3(xdiv(h.x, h.w), xdiv(h.y, h.w), xdiv(h.z, h.w));
        }
        const f3 sun_dir = mk3(p.uniforms.sun_dir.x, p.uniforms.sun_dir.y, p.uniforms.sun_dir.z);
        if (accel_occluded(p.accel, pos, sun_dir, kSunTMax)) m[4] = 1u;
        const uint32_t cluster = cluster_index(fx, fy, depth, p.uniforms);
        if (cluster < p.n_clusters) {
            const uint32_t count = min(__ldg(p.cluster_counts + cluster), TR_MAX_LIGHTS_PER_CLUSTER);
            const uint32_t* list = p.cluster_indices + (size_t)cluster * TR_MAX_LIGHTS_PER_CLUSTER;
            for (uint32_t i = 0; i < count; i++) {
                const float4 lp = __ldg(reinterpret_cast<const float4*>(p.lights + __ldg(list + i)));
                const f3 vec = xsub3(mk3(lp.x, lp.y, lp.z), pos);
                const float dist = xsqrt(xdot3(vec, vec));
                if (accel_occluded(p.accel, pos, xdivs3(vec, dist), dist)) m[i >> 5] |= 1u << (i & 31u);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) p.shadow_mask[(size_t)p.shadow_plane * k + g] = m[k];
}
#endif

namespace tr {
#if TR_SHADE_SHADOW
int32_t launch_shade_opaque_shadowed(const ShadeLaunch& p, int sm_count, cudaStream_t s) { return launch_any<false>(p, sm_count, s); }
int32_t launch_shade_transmission_shadowed(const ShadeLaunch& p, int sm_count, cudaStream_t s) { return launch_any<true>(p, sm_count, s); }
int32_t launch_shadow_mask(const ShadeLaunch& p, int sm_count, cudaStream_t s) {
    const uint32_t n_px = p.px_end - p.px_begin;
    if (n_px == 0) return TR_OK;
    if (!p.shadow_mask) return fail(TR_ERR_STATE, "shadow pass without a mask buffer");
    if (p.position) shadow_mask_kernel<true><<<(n_px + 127) / 128, 128, 0, s>>>(p);
    else shadow_mask_kernel<false><<<(n_px + 127) / 128, 128, 0, s>>>(p);
    count_launches(1);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}
#else
int32_t launch_shade_opaque(const ShadeLaunch& p, int sm_count, cudaStream_t s) { return launch_any<false>(p, sm_count, s); }
int32_t launch_shade_transmission(const ShadeLaunch& p, int sm_count, cudaStream_t s) { return launch_any<true>(p, sm_count, s); }
#endif
}  // namespace tr

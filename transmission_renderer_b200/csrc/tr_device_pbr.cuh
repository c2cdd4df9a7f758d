// tr_device_pbr.cuh — the glam-pbr function contracts as sm_100a device code.
//
// Contracts kept (reference: /root/reference/glam-pbr/src/lib.rs):
//   basic_brdf                :377-423     transmission_btdf       :200-233
//   ibl_volume_refraction     :292-354     light_direction_and_attenuation :12-23
//   d_ggx :101-109  v_smith_ggx_correlated :114-133  fresnel_schlick :137-139
//   calculate_combined_f0/f90 :425-435     IndexOfRefraction::to_dielectric_f0 :192-195
// Re-designed for the GPU: everything that does not depend on the light is
// hoisted into a per-pixel `PixelShading` record (built once, reused for the
// sun and every clustered light), the n.h chain runs in the exact regime and the
// rest in the fast regime (see tr_device_math.cuh).
#pragma once

#include "tr_device_math.cuh"

namespace trd {

#define TR_F32_EPSILON 1.1920929e-7f
#define TR_PI 3.14159265358979323846f
#define TR_FRAC_1_PI 0.318309886183790671538f

struct MaterialParams {  // glam-pbr lib.rs:171-179
    f3 diffuse_colour;
    float metallic;
    float perceptual_roughness;
    float index_of_refraction;
    f3 specular_colour;
    float specular_factor;
};

struct BrdfResult {  // glam-pbr lib.rs:437-441
    f3 diffuse, specular;
};

TRD float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// IndexOfRefraction::to_dielectric_f0, lib.rs:192-195
TRD float to_dielectric_f0(float ior) {
    float root = (ior - 1.0f) * frcp(ior + 1.0f);
    return root * root;
}

// Light-independent part of basic_brdf / transmission_btdf for one pixel.
struct PixelShading {
    f3 n, v;            // unit normal / view (exact regime)
    float nov;          // Dot::new(normal, view), clamped (lib.rs:92-99)
    float nov_raw;      // n.v before the clamp (the fast light loop derives n.h and v.l' from raw dot products)
    f3 f0, df;          // combined f0 and (f90 - f0)           (lib.rs:425-435)
    f3 c_diff_pi;       // lerp(base, 0, metallic) / pi         (lib.rs:404, 359)
    f3 base;            // diffuse_colour
    float a2, a2m1;     // alpha^2 and alpha^2 - 1 (exact), alpha = r^2  (lib.rs:104-106)
    float one_m_a2;     // 1 - alpha^2                                   (lib.rs:122-124)
    float nov2_term;    // nov^2 (1 - a2) + a2
    // transmission lobe: alpha_t = alpha * clamp(2 ior - 2, 0, 1)       (lib.rs:144-148, 209)
    float at2, at2m1, one_m_at2, nov2_term_t;
    // hoisted for the fast light loop: sqrt(nov2_term[_t]) and alpha^2 / (2 pi)
    float s_nov, s_nov_t, a2_2pi, at2_2pi;
};

// calculate_combined_f0 / f90 (lib.rs:425-435) and c_diff / pi (lib.rs:404, 359): the colour factors of a pixel
TRD void colour_factors(const MaterialParams& m, f3& f0, f3& df, f3& c_diff_pi) {
    float d0 = to_dielectric_f0(m.index_of_refraction);
    f3 dielectric = scale3(scale3(m.specular_colour, d0), m.specular_factor);
    f0 = lerp3(dielectric, m.diffuse_colour, m.metallic);
    f3 f90 = lerp3(splat3(m.specular_factor), splat3(1.0f), m.metallic);
    df = sub3(f90, f0);
    c_diff_pi = scale3(lerp3(m.diffuse_colour, splat3(0.0f), m.metallic), TR_FRAC_1_PI);
}

TRD PixelShading make_pixel_shading(const MaterialParams& m, f3 n, f3 v, bool with_transmission) {
    PixelShading s;
    s.n = n;
    s.v = v;
    s.nov_raw = xdot3(n, v);
    s.nov = fmaxf(s.nov_raw, TR_F32_EPSILON);
    colour_factors(m, s.f0, s.df, s.c_diff_pi);
    s.base = m.diffuse_colour;
    float alpha = xmul(m.perceptual_roughness, m.perceptual_roughness);
    s.a2 = xmul(alpha, alpha);
    s.a2m1 = xsub(s.a2, 1.0f);
    s.one_m_a2 = 1.0f - s.a2;
    s.nov2_term = fmaf(s.nov * s.nov, s.one_m_a2, s.a2);
    s.s_nov = fsqrt(s.nov2_term);
    s.a2_2pi = s.a2 * (0.5f * TR_FRAC_1_PI);
    s.s_nov_t = s.at2_2pi = 0.0f;
    if (with_transmission) {
        float c = xsub(xmul(m.index_of_refraction, 2.0f), 2.0f);
        c = fminf(fmaxf(c, 0.0f), 1.0f);
        float at = xmul(alpha, c);
        s.at2 = xmul(at, at);
        s.at2m1 = xsub(s.at2, 1.0f);
        s.one_m_at2 = 1.0f - s.at2;
        s.nov2_term_t = fmaf(s.nov * s.nov, s.one_m_at2, s.at2);
        s.s_nov_t = fsqrt(s.nov2_term_t);
        s.at2_2pi = s.at2 * (0.5f * TR_FRAC_1_PI);
    } else {
        s.at2 = s.at2m1 = s.one_m_at2 = s.nov2_term_t = 0.0f;
    }
    return s;
}

// d_ggx * v_smith_ggx_correlated for one (noh, nol) — lib.rs:101-133.
// `noh` must come from the exact chain; f is evaluated exactly, the rest fast.
TRD float ggx_d_times_v(float noh, float nol, float nov, float a2, float a2m1, float one_m_a2, float nov2_term) {
    float f = xadd(xmul(xmul(noh, noh), a2m1), 1.0f);
    float d = a2 * frcp(TR_PI * f * f);
    float ggx_v = nol * fsqrt(nov2_term);
    float ggx_l = nov * fsqrt(fmaf(nol * nol, one_m_a2, a2));
    float ggx = ggx_v + ggx_l;
    float vis = ggx > 0.0f ? 0.5f * frcp(ggx) : 0.0f;
    return d * vis;
}

// fresnel_schlick, lib.rs:137-139: f0 + (f90 - f0) * (1 - v.h)^5
TRD f3 fresnel_schlick(float voh, f3 f0, f3 df) {
    float x = 1.0f - voh;
    float x2 = x * x;
    float p = x2 * x2 * x;
    return fma3(df, p, f0);
}

// One light through basic_brdf (lib.rs:377-423).  `l` is the unit light direction from the exact chain.
TRD void brdf_light(const PixelShading& s, f3 l, f3 light_intensity, f3& diffuse_acc, f3& specular_acc) {
    f3 h = xnormalize3(xadd3(s.v, l));                          // Halfway::new, lib.rs:64-68
    float noh = fmaxf(xdot3(s.n, h), TR_F32_EPSILON);
    // the clamped dots stay exact too: near the silhouette (n.l, n.v -> 0) the visibility term divides by them
    float nol = fmaxf(xdot3(s.n, l), TR_F32_EPSILON);
    float voh = fmaxf(xdot3(s.v, h), TR_F32_EPSILON);
    f3 fresnel = fresnel_schlick(voh, s.f0, s.df);
    f3 li = scale3(light_intensity, nol);
    float kd = 1.0f - max_element3(fresnel);                    // diffuse_brdf, lib.rs:356-360
    float dv = ggx_d_times_v(noh, nol, s.nov, s.a2, s.a2m1, s.one_m_a2, s.nov2_term);
    diffuse_acc = add3(diffuse_acc, mul3(li, scale3(s.c_diff_pi, kd)));
    specular_acc = add3(specular_acc, mul3(li, scale3(fresnel, dv)));
}

// One light through transmission_btdf (lib.rs:200-233), unweighted.
TRD f3 btdf_light(const PixelShading& s, f3 l) {
    // light + 2 n dot(-light, n), lib.rs:211
    f3 lm = xnormalize3(xadd3(l, xscale3(xscale3(s.n, 2.0f), -xdot3(l, s.n))));
    f3 h = xnormalize3(xadd3(s.v, lm));
    float noh = fmaxf(xdot3(s.n, h), TR_F32_EPSILON);
    float voh = fmaxf(xdot3(s.v, h), TR_F32_EPSILON);
    float nolm = fmaxf(xdot3(s.n, lm), TR_F32_EPSILON);
    float dv = ggx_d_times_v(noh, nolm, s.nov, s.at2, s.at2m1, s.one_m_at2, s.nov2_term_t);
    f3 fresnel = fresnel_schlick(voh, s.f0, s.df);
    f3 one_m_f = mk3(1.0f - fresnel.x, 1.0f - fresnel.y, 1.0f - fresnel.z);
    return mul3(scale3(one_m_f, dv), s.base);
}

// ---- clustered point lights: adaptive exactness -------------------------------------------------------
// The only ill-conditioned step of a light evaluation is f = noh^2 (a^2 - 1) + 1 (d_ggx, lib.rs:101-109): one ulp of
// n.h moves D by 4 ulp / f.  Everything else is well conditioned, so the light loop runs in the fast regime and
// re-derives n.h through the exact chain (position -> direction -> halfway -> n.h, the oracle's operation order)
// only when the fast estimate of f is below TR_EXACT_F — a few percent of the (pixel, light) pairs, all of them inside
// highlights.  There the lobe is bit-identical to the all-exact evaluation; elsewhere it is within ~1e-6 of it.  Measured on
// B200 with the round-2 loop (un-normalised halfway vector) at the 0.04 below: worst rel-L2 of a final fp32 frame against the
// oracle 7.5e-6 (tests/test_gpu_parity.py spheres cases), worst element of tests/test_gpu_hot_loop.py 0.59 of its tolerance.
// (Round 1's loop, which took |h|^2 from 2 (1 + v.l), needed 0.08: 0 (all fast) 1.0e-3, 0.02 1.9e-4, 0.04 6.7e-5, 0.08 1.2e-5.)
#ifndef TR_EXACT_F
#define TR_EXACT_F 0.04f
#endif

TRD f3 exact_light_dir(f3 vec) {  // light_direction_and_attenuation, lib.rs:12-23, exact regime
    return xunit3_mid(vec);
}

// light_direction_and_attenuation, lib.rs:12-23 (direction exact, attenuation fast)
TRD void light_direction_and_attenuation(f3 fragment_position, f3 light_position, f3& direction, float& attenuation) {
    f3 vec = xsub3(light_position, fragment_position);
    float d2 = xdot3(vec, vec);
    float dist = xsqrt(d2);
    direction = xdivs3(vec, dist);
    attenuation = frcp(d2);
}

TRD float exact_noh(f3 n, f3 v, f3 l) {
    f3 h = xnormalize3_mid(xadd3(v, l));                             // Halfway::new, lib.rs:64-68
    return fmaxf(xdot3(n, h), TR_F32_EPSILON);
}

// sm_100a retires two fp32 FMAs per issue slot as one packed instruction (PTX fma.rn.f32x2, SASS FFMA2) when both halves
// live in an aligned register pair.  The accumulators are kept as such pairs — (d, s0) and (s1, t1) per colour channel —
// and the light table stores every colour channel twice, so one FFMA2 adds (colour, colour) * (w_a, w_b) to both sums.
typedef unsigned long long f32x2;
TRD f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
TRD void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
TRD void ffma2(f32x2& acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
TRD f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
TRD f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
TRD f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
TRD f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// ---- the light loop's fast regime ("lean" form) -----------------------------------------------------------
// Everything the loop needs of a pixel, and nothing else (registers are what bounds the kernels' occupancy):
struct LoopPixel {
    f3 pos, n, v;                           // world position, unit normal, unit view vector
    f32x2 pos_xy, n_xy, v_xy;               // their x and y as packed pairs: the loop's vector adds / scales take x and y in one
                                            // instruction (FADD2 / FMUL2 / FFMA2); the same registers as above, no copies
    float nov_raw, nov;                     // n.v, and Dot::new's clamp of it (lib.rs:92-99)
    float a2, a2m1, one_m_a2, s_nov;        // reflection lobe: alpha^2, alpha^2 - 1, 1 - alpha^2, sqrt(nov^2 (1 - a2) + a2)
    float omf0m, ndfm;                      // 1 - f0[m] and -(f90 - f0)[m] / 32, m = the channel with the largest f0 (see below)
    // transmission lobe (alpha_t = alpha clamp(2 ior - 2, 0, 1), lib.rs:144-148,209)
    float at2, at2m1, one_m_at2, s_nov_t;
};
// Sums over the lights of a pixel.  The per-pixel colour factors are applied once, after the loop (finish_sums):
//   d  = sum li * (1 - max F)                  diffuse  = c_diff / pi * d                  (lib.rs:356-360, 404)
//   s0 = sum li * D V, s1 = sum li * D V * p   specular = f0 * s0 + (f90 - f0) * s1        (F = f0 + (f90 - f0) p, lib.rs:137-139)
//   t0 = sum l * D' V', t1 = sum l * D' V' p'  transmission = base * ((1 - f0) t0 - (f90 - f0) t1)   (lib.rs:226-232)
// with li = light intensity * n.l, p = (1 - v.h)^5 (carried as 32 p, see light_lean), and D V without its
// alpha^2 / (2 pi) (a per-pixel constant too).
// max_element(F) (lib.rs:359): f90 = lerp(splat(specular_factor), 1, metallic) has three equal channels (lib.rs:432-435), so
// the three lines f0_c + (f90 - f0_c) p meet at p = 1 and the channel with the largest f0 is the largest on all of [0, 1].
#define TR_2_POW_M5 0.03125f

TRD uint32_t argmax3(f3 a) { return a.x >= a.y ? (a.x >= a.z ? 0u : 2u) : (a.y >= a.z ? 1u : 2u); }
TRD float pick3(f3 a, uint32_t i) { return i == 0u ? a.x : (i == 1u ? a.y : a.z); }

TRD LoopPixel make_loop_pixel(const PixelShading& s, f3 pos) {
    LoopPixel q;
    q.pos = pos; q.n = s.n; q.v = s.v;
    q.pos_xy = pack2(pos.x, pos.y); q.n_xy = pack2(s.n.x, s.n.y); q.v_xy = pack2(s.v.x, s.v.y);
    q.nov_raw = s.nov_raw; q.nov = s.nov;
    q.a2 = s.a2; q.a2m1 = s.a2m1; q.one_m_a2 = s.one_m_a2; q.s_nov = s.s_nov;
    const uint32_t m = argmax3(s.f0);
    q.omf0m = 1.0f - pick3(s.f0, m);
    q.ndfm = -pick3(s.df, m) * TR_2_POW_M5;
    q.at2 = s.at2; q.at2m1 = s.at2m1; q.one_m_at2 = s.one_m_at2; q.s_nov_t = s.s_nov_t;
    return q;
}

struct Colour2 { f32x2 r, g, b; };   // (r, r), (g, g), (b, b)
TRD Colour2 dup_colour(f3 c) { Colour2 k; k.r = pack2(c.x, c.x); k.g = pack2(c.y, c.y); k.b = pack2(c.z, c.z); return k; }
TRD float lo2(f32x2 v) { float a, b; unpack2(v, a, b); return a; }

struct LoopSums {
    f32x2 ds_r, ds_g, ds_b;   // (d, s0) per channel
    f32x2 st_r, st_g, st_b;   // (s1, t1) per channel (transmissive pass)
    f3 s1;                    // s1 alone (opaque pass: no t1 to pair it with)
    f3 t0;
    TRD void clear() { ds_r = ds_g = ds_b = st_r = st_g = st_b = 0ull; t0.x = t0.y = t0.z = 0.0f; s1 = t0; }
};
// One light: basic_brdf (lib.rs:377-423) and, with TRANS, transmission_btdf (lib.rs:200-233) for the same light.
//   l = the unit light direction, nol_raw = n.l; colour * factor = the light's intensity at the fragment (emission x
//   attenuation x spotlight factor); exact_dir() = l of the exact chain (only evaluated inside highlights).
// The halfway vector is formed un-normalised, h = v + l, and only its squared length is used: n.h = (n.v + n.l) / |h| and,
// for unit v and l, v.h = |h| / 2 — so y = 2 - |h| = 2 (1 - v.h) and y^5 = 32 (1 - v.h)^5.  (|h|^2 = 2 (1 + v.l) would save
// three instructions, but 1 + v.l cancels when the light is almost behind the viewer, the grazing forward-scatter
// configuration: tests/test_gpu_hot_loop.py.)  The mirrored light of the BTDF, l' = l - 2 (n.l) n, is a reflection:
// h' = v + l' = h - 2 (n.l) n and n.l' = -n.l.  D V = (a^2 / 2 pi) / (f^2 ggx) with f = noh^2 (a^2 - 1) + 1 (lib.rs:101-133 in
// one reciprocal; ggx > 0 because n.v and n.l are clamped to EPSILON).  f is ill-conditioned inside a highlight: below
// TR_EXACT_F it is re-derived through the exact chain.  The transmission lobe has no n.l' factor and grows like 1 / ggx, so
// where ggx is below TR_EXACT_GGX (grazing light on a pixel whose n.v is clamped) n.l' comes from the exact chain too.
// Both lobes share one (rarely taken) branch.
#ifndef TR_EXACT_GGX
#define TR_EXACT_GGX 0.005f
#endif
template <bool TRANS, typename ExactDir>
TRD void light_lean(const LoopPixel& s, ExactDir exact_dir, f3 l, float nol_raw, const Colour2& colour, float factor, LoopSums& a) {
    const float nol = fmaxf(nol_raw, TR_F32_EPSILON);
    const f32x2 h_xy = add2(s.v_xy, pack2(l.x, l.y));
    f3 h;
    unpack2(h_xy, h.x, h.y);
    h.z = s.v.z + l.z;
    const float h2 = fmaxf(dot3(h, h), 1e-30f);
    const float r = frsqrt(h2);
    const float noh = fmaxf((s.nov_raw + nol_raw) * r, TR_F32_EPSILON);
    float f = fmaf(noh * noh, s.a2m1, 1.0f);
    const float y = 2.0f - h2 * r;
    const float ggx = fmaf(nol, s.s_nov, s.nov * fsqrt(fmaf(nol * nol, s.one_m_a2, s.a2)));
    bool slow = f < TR_EXACT_F;
    // transmission lobe
    float nolt = 0.0f, ft = 1.0f, yt = 0.0f, ggxt = 1.0f;
    if (TRANS) {
        nolt = fmaxf(-nol_raw, TR_F32_EPSILON);
        const float k2 = -2.0f * nol_raw;
        f3 ht;
        unpack2(fma2(s.n_xy, pack2(k2, k2), h_xy), ht.x, ht.y);
        ht.z = fmaf(s.n.z, k2, h.z);
        const float ht2 = fmaxf(dot3(ht, ht), 1e-30f);
        const float rt = frsqrt(ht2);
        const float noht = fmaxf((s.nov_raw - nol_raw) * rt, TR_F32_EPSILON);
        ft = fmaf(noht * noht, s.at2m1, 1.0f);
        yt = 2.0f - ht2 * rt;
        ggxt = fmaf(nolt, s.s_nov_t, s.nov * fsqrt(fmaf(nolt * nolt, s.one_m_at2, s.at2)));
        slow = slow || ft < TR_EXACT_F || ggxt < TR_EXACT_GGX;
    }
    if (slow) {
        const f3 lx = exact_dir();
        if (f < TR_EXACT_F) {
            const float e = exact_noh(s.n, s.v, lx);
            f = xadd(xmul(xmul(e, e), s.a2m1), 1.0f);
        }
        if (TRANS && (ft < TR_EXACT_F || ggxt < TR_EXACT_GGX)) {
            const f3 lmx = xnormalize3_mid(xadd3(lx, xscale3(xscale3(s.n, 2.0f), -xdot3(lx, s.n))));   // lib.rs:211
            const float e = exact_noh(s.n, s.v, lmx);
            const float nolx = fmaxf(xdot3(s.n, lmx), TR_F32_EPSILON);
            ft = xadd(xmul(xmul(e, e), s.at2m1), 1.0f);
            ggxt = fmaf(nolx, s.s_nov_t, s.nov * fsqrt(fmaf(nolx * nolx, s.one_m_at2, s.at2)));
        }
    }
    const float y2 = y * y;
    const float p = y2 * y2 * y;                                            // 32 x fresnel_schlick's powf(1 - v.h, 5)
    const float w = factor * nol;
    const float wdv = w * frcp(f * f * ggx);
    const f32x2 w_ds = pack2(w * fmaf(s.ndfm, p, s.omf0m), wdv);
    ffma2(a.ds_r, colour.r, w_ds);
    ffma2(a.ds_g, colour.g, w_ds);
    ffma2(a.ds_b, colour.b, w_ds);
    const f3 c1 = mk3(lo2(colour.r), lo2(colour.g), lo2(colour.b));
    if (TRANS) {
        const float yt2 = yt * yt;
        const float pt = yt2 * yt2 * yt;
        const float wt = factor * frcp(ft * ft * ggxt);
        a.t0 = fma3(c1, wt, a.t0);
        const f32x2 w_st = pack2(wdv * p, wt * pt);
        ffma2(a.st_r, colour.r, w_st);
        ffma2(a.st_g, colour.g, w_st);
        ffma2(a.st_b, colour.b, w_st);
    } else {
        a.s1 = fma3(c1, wdv * p, a.s1);
    }
}

// A point light of the clustered loops (lighting.rs:58-92, 179-216): light_direction_and_attenuation (lib.rs:12-23) in the
// fast regime, then light_lean.  The direction is formed first and the dot products taken with it, as the exact chain does:
// for a grazing light n.l cancels, and only then does its rounding follow the reference's operation order.
// `spot(dir, factor)` may scale the factor (Light::spotlight_factor, opaque loop only).
template <bool TRANS, typename Spot>
TRD void point_light_lean(const LoopPixel& lp, f3 light_position, const Colour2& colour, Spot spot, LoopSums& sums) {
    const f32x2 vec_xy = sub2(pack2(light_position.x, light_position.y), lp.pos_xy);
    f3 vec, dir;
    unpack2(vec_xy, vec.x, vec.y);
    vec.z = light_position.z - lp.pos.z;
    const float inv_d = frsqrt(dot3(vec, vec));
    unpack2(mul2(vec_xy, pack2(inv_d, inv_d)), dir.x, dir.y);
    dir.z = vec.z * inv_d;
    float factor = inv_d * inv_d;
    spot(dir, factor);
    light_lean<TRANS>(lp, [&]() { return exact_light_dir(vec); }, dir, dot3(lp.n, dir), colour, factor, sums);
}

// The per-pixel colour factors, once for all lights of the pixel (see LoopSums).
TRD void finish_sums(const LoopSums& a, f3 f0, f3 df, f3 c_diff_pi, f3 base, float a2, float at2, bool with_transmission,
                     f3& diffuse, f3& specular, f3& transmission) {
    const float k = a2 * (0.5f * TR_FRAC_1_PI);
    const f3 dfs = scale3(df, TR_2_POW_M5);   // s1 / t1 carry 32 p
    f3 d, s0, s1, t1;
    unpack2(a.ds_r, d.x, s0.x); unpack2(a.ds_g, d.y, s0.y); unpack2(a.ds_b, d.z, s0.z);
    unpack2(a.st_r, s1.x, t1.x); unpack2(a.st_g, s1.y, t1.y); unpack2(a.st_b, s1.z, t1.z);
    if (!with_transmission) s1 = a.s1;
    diffuse = mul3(d, c_diff_pi);
    specular = scale3(fma3v(f0, s0, mul3(dfs, s1)), k);
    transmission = splat3(0.0f);
    if (with_transmission) {
        const float kt = at2 * (0.5f * TR_FRAC_1_PI);
        const f3 omf0 = mk3(1.0f - f0.x, 1.0f - f0.y, 1.0f - f0.z);
        const f3 t = sub3(mul3(omf0, a.t0), mul3(dfs, t1));
        transmission = scale3(mul3(t, base), kt);
    }
}

// ---- contract-shaped wrappers (used by the tr_eval_* batch evaluators) ----
struct BasicBrdfParams {  // lib.rs:163-169
    f3 normal, light, light_intensity, view;
    MaterialParams material_params;
};

TRD BrdfResult basic_brdf(const BasicBrdfParams& p) {
    PixelShading s = make_pixel_shading(p.material_params, p.normal, p.view, false);
    BrdfResult r;
    r.diffuse = splat3(0.0f);
    r.specular = splat3(0.0f);
    brdf_light(s, p.light, p.light_intensity, r.diffuse, r.specular);
    return r;
}

TRD f3 transmission_btdf(const MaterialParams& m, f3 normal, f3 view, f3 light) {
    PixelShading s = make_pixel_shading(m, normal, view, true);
    return btdf_light(s, light);
}

// ---- sampled images (oracle: oracle/shade.c) ----
struct PyramidDesc {
    const uint2* base;      // RGBA16F texels, all levels in one allocation
    uint32_t levels;
    uint32_t w[16], h[16];
    uint32_t offset[16];    // texel offset of each level
};
struct LutDesc {
    const uchar2* rg;       // R,G of the RGBA8 UNORM LUT
    uint32_t w, h;
};

TRD void bilinear_setup(float u, uint32_t size, uint32_t& i0, uint32_t& i1, float& frac) {
    float p = xsub(xmul(u, (float)size), 0.5f);
    if (!(p == p)) p = 0.0f;
    float fl = floorf(p);
    float f = xsub(p, fl);
    int i;
    if (fl <= -2147483648.0f) { i = INT_MIN; f = 0.0f; }
    else if (fl >= 2147483520.0f) { i = 2147483520; f = 0.0f; }
    else i = (int)fl;
    int hi = (int)size - 1;
    i0 = (uint32_t)min(max(i, 0), hi);
    int ip1 = i >= 2147483520 ? i : i + 1;
    i1 = (uint32_t)min(max(ip1, 0), hi);
    frac = f;
}

TRD f3 sample_level(const PyramidDesc& p, uint32_t level, float u, float v) {
    uint32_t w = p.w[level], h = p.h[level];
    const uint2* d = p.base + p.offset[level];
    uint32_t x0, x1, y0, y1;
    float fx, fy;
    bilinear_setup(u, w, x0, x1, fx);
    bilinear_setup(v, h, y0, y1, fy);
    f4 t00 = unpack_rgba16f(__ldg(d + (size_t)y0 * w + x0));
    f4 t10 = unpack_rgba16f(__ldg(d + (size_t)y0 * w + x1));
    f4 t01 = unpack_rgba16f(__ldg(d + (size_t)y1 * w + x0));
    f4 t11 = unpack_rgba16f(__ldg(d + (size_t)y1 * w + x1));
    f3 top = lerp3(mk3(t00.x, t00.y, t00.z), mk3(t10.x, t10.y, t10.z), fx);
    f3 bot = lerp3(mk3(t01.x, t01.y, t01.z), mk3(t11.x, t11.y, t11.z), fx);
    return lerp3(top, bot, fy);
}

// framebuffer.sample_by_lod(clamp_sampler, uv, lod).rgb — shader/src/lib.rs:135-138
TRD f3 sample_pyramid(const PyramidDesc& p, float u, float v, float lod) {
    float max_lod = (float)(p.levels - 1);
    if (!(lod > 0.0f)) lod = 0.0f;
    if (lod > max_lod) lod = max_lod;
    float l0f = floorf(lod);
    uint32_t l0 = (uint32_t)l0f;
    uint32_t l1 = l0 + 1 < p.levels ? l0 + 1 : p.levels - 1;
    float t = lod - l0f;
    f3 s0 = sample_level(p, l0, u, v);
    f3 s1 = sample_level(p, l1, u, v);
    return lerp3(s0, s1, t);
}

// textures[ggx_lut].sample(clamp_sampler, (n.v, roughness)).xy — shader/src/lib.rs:126-133
TRD float2 sample_lut(const LutDesc& lut, float nov, float roughness) {
    uint32_t x0, x1, y0, y1;
    float fx, fy;
    bilinear_setup(nov, lut.w, x0, x1, fx);
    bilinear_setup(roughness, lut.h, y0, y1, fy);
    uchar2 t00 = __ldg(lut.rg + (size_t)y0 * lut.w + x0);
    uchar2 t10 = __ldg(lut.rg + (size_t)y0 * lut.w + x1);
    uchar2 t01 = __ldg(lut.rg + (size_t)y1 * lut.w + x0);
    uchar2 t11 = __ldg(lut.rg + (size_t)y1 * lut.w + x1);
    const float k = 1.0f / 255.0f;
    float ax = fmaf((float)t10.x * k - (float)t00.x * k, fx, (float)t00.x * k);
    float bx = fmaf((float)t11.x * k - (float)t01.x * k, fx, (float)t01.x * k);
    float ay = fmaf((float)t10.y * k - (float)t00.y * k, fx, (float)t00.y * k);
    float by = fmaf((float)t11.y * k - (float)t01.y * k, fx, (float)t01.y * k);
    return make_float2(fmaf(bx - ax, fy, ax), fmaf(by - ay, fy, ay));
}

// ---- material textures (row N2): texture.sample(sampler, uv) with the repeat sampler, shader/src/lib.rs:251-267.
// Everything here runs in the exact regime and mirrors oracle/shade.c line by line: the sampled values feed the shading
// normal and the roughness, i.e. the ill-conditioned n.h chain, and the level of detail is a discrete decision.
struct TexDesc {
    const float4* base;     // decoded texels (sRGB -> linear / UNORM -> [0,1] done at upload), all levels
    uint32_t w, h, levels, srgb;
    uint32_t off[16];       // texel offset of each level
};

TRD float xlerp1(float a, float b, float t) { return xadd(a, xmul(xsub(b, a), t)); }

TRD void repeat_setup(float u, uint32_t size, uint32_t& i0, uint32_t& i1, float& frac) {
    float p = xsub(xmul(u, (float)size), 0.5f);
    if (!(p == p) || !(fabsf(p) < 1.0e9f)) p = 0.0f;
    const float fl = floorf(p);
    const int i = (int)fl, n = (int)size;
    int a = i % n, b = (i + 1) % n;
    i0 = (uint32_t)(a < 0 ? a + n : a);
    i1 = (uint32_t)(b < 0 ? b + n : b);
    frac = xsub(p, fl);
}

TRD f4 sample_texture_level(const TexDesc& t, uint32_t level, float u, float v) {
    uint32_t w = t.w >> level, h = t.h >> level;
    if (w == 0) w = 1;
    if (h == 0) h = 1;
    const float4* d = t.base + t.off[level];
    uint32_t x0, x1, y0, y1;
    float fx, fy;
    repeat_setup(u, w, x0, x1, fx);
    repeat_setup(v, h, y0, y1, fy);
    const float4 t00 = __ldg(d + (size_t)y0 * w + x0), t10 = __ldg(d + (size_t)y0 * w + x1);
    const float4 t01 = __ldg(d + (size_t)y1 * w + x0), t11 = __ldg(d + (size_t)y1 * w + x1);
    f4 r;
    r.x = xlerp1(xlerp1(t00.x, t10.x, fx), xlerp1(t01.x, t11.x, fx), fy);
    r.y = xlerp1(xlerp1(t00.y, t10.y, fx), xlerp1(t01.y, t11.y, fx), fy);
    r.z = xlerp1(xlerp1(t00.z, t10.z, fx), xlerp1(t01.z, t11.z, fx), fy);
    r.w = xlerp1(xlerp1(t00.w, t10.w, fx), xlerp1(t01.w, t11.w, fx), fy);
    return r;
}

// duv = (du/dx, dv/dx, du/dy, dv/dy): differences to the right / lower neighbour (Vulkan level-of-detail operation)
TRD f4 sample_texture(const TexDesc& t, float u, float v, float4 duv) {
    const float ux = xmul(duv.x, (float)t.w), vx = xmul(duv.y, (float)t.h);
    const float uy = xmul(duv.z, (float)t.w), vy = xmul(duv.w, (float)t.h);
    const float rho2 = rmax(xadd(xmul(ux, ux), xmul(vx, vx)), xadd(xmul(uy, uy), xmul(vy, vy)));
    float lod = xmul(0.5f, xlog2_spec(rho2));
    const float max_lod = (float)(t.levels - 1);
    if (!(lod > 0.0f)) lod = 0.0f;
    if (lod > max_lod) lod = max_lod;
    const float l0f = floorf(lod);
    const uint32_t l0 = (uint32_t)l0f;
    const uint32_t l1 = l0 + 1 < t.levels ? l0 + 1 : t.levels - 1;
    const float f = xsub(lod, l0f);
    const f4 a = sample_texture_level(t, l0, u, v), b = sample_texture_level(t, l1, u, v);
    f4 r;
    r.x = xlerp1(a.x, b.x, f);
    r.y = xlerp1(a.y, b.y, f);
    r.z = xlerp1(a.z, b.z, f);
    r.w = xlerp1(a.w, b.w, f);
    return r;
}

struct TextureSampler {  // shader/src/lib.rs:251-267
    const TexDesc* textures;
    uint32_t n_textures;
    float u, v;
    float4 duv;
    TRD f4 sample(int32_t id) const {
        f4 z;
        z.x = z.y = z.z = z.w = 0.0f;
        if ((uint32_t)id >= n_textures || textures[id].base == nullptr) return z;  // robust access: unbound image reads 0
        return sample_texture(textures[id], u, v, duv);
    }
};

// calculate_normal + compute_cotangent_frame, shader/src/lighting.rs:222-259 (oracle/shade.c calculate_normal)
TRD f3 calculate_normal(f3 interpolated, int32_t normal_map, const TextureSampler& ts, f3 dpos_dx, f3 dpos_dy) {
    f3 normal = xnormalize3(interpolated);
    if (normal_map != -1) {
        const f4 smp = ts.sample(normal_map);
        const float k = xdiv(128.0f, 127.0f);
        const f3 m = mk3(xsub(xdiv(xmul(smp.x, 255.0f), 127.0f), k), xsub(xdiv(xmul(smp.y, 255.0f), 127.0f), k),
                         xsub(xdiv(xmul(smp.z, 255.0f), 127.0f), k));
        const f3 dp2perp = xcross3(dpos_dy, normal), dp1perp = xcross3(normal, dpos_dx);
        const f3 t = xadd3(xscale3(dp2perp, ts.duv.x), xscale3(dp1perp, ts.duv.z));
        const f3 b = xadd3(xscale3(dp2perp, ts.duv.y), xscale3(dp1perp, ts.duv.w));
        const float invmax = xdiv(1.0f, xsqrt(rmax(xdot3(t, t), xdot3(b, b))));
        const f3 c0 = xscale3(t, invmax), c1 = xscale3(b, invmax);
        normal = xnormalize3(xadd3(xadd3(xscale3(c0, m.x), xscale3(c1, m.y)), xscale3(normal, m.z)));
    }
    return normal;
}

struct IblVolumeRefractionParams {  // lib.rs:235-246; proj_view and log2(size_x) are per-launch constants
    MaterialParams material_params;
    f3 normal, view, position;
    float thickness, model_scale, attenuation_distance;
    f3 attenuation_colour;
};

// ibl_volume_refraction, lib.rs:292-354.  The FSamp/GSamp closures of the reference are the
// pyramid / LUT descriptors here.  `f0`/`df` may be passed pre-computed by the caller.
TRD f3 ibl_volume_refraction(const IblVolumeRefractionParams& p, const mat4& proj_view, float log2_size_x,
                             const PyramidDesc& fb, const LutDesc& lut, f3 f0, f3 df) {
    const MaterialParams& m = p.material_params;
    // refract(-view, normal, ior), lib.rs:248-256
    float eta = frcp(m.index_of_refraction);
    f3 incident = neg3(p.view);
    float ndi = dot3(p.normal, incident);
    float k = 1.0f - eta * eta * (1.0f - ndi * ndi);
    float t = fmaf(eta, ndi, fsqrt(k));
    f3 refr = sub3(scale3(incident, eta), scale3(p.normal, t));
    // get_volume_transmission_ray, lib.rs:258-268
    float ray_length = p.thickness * p.model_scale;
    f3 ray = scale3(normalize3(refr), ray_length);
    f3 exit_p = add3(p.position, ray);
    // project, lib.rs:330-332
    float dx = fmaf(proj_view.c[3][0], 1.0f, fmaf(proj_view.c[2][0], exit_p.z, fmaf(proj_view.c[1][0], exit_p.y, proj_view.c[0][0] * exit_p.x)));
    float dy = fmaf(proj_view.c[3][1], 1.0f, fmaf(proj_view.c[2][1], exit_p.z, fmaf(proj_view.c[1][1], exit_p.y, proj_view.c[0][1] * exit_p.x)));
    float dw = fmaf(proj_view.c[3][3], 1.0f, fmaf(proj_view.c[2][3], exit_p.z, fmaf(proj_view.c[1][3], exit_p.y, proj_view.c[0][3] * exit_p.x)));
    float inv_w = 1.0f / dw;
    float tu = (dx * inv_w + 1.0f) * 0.5f;
    float tv = (dy * inv_w + 1.0f) * 0.5f;
    // lod, lib.rs:334-335 (PerceptualRoughness::apply_ior :158-160)
    float lod = log2_size_x * (m.perceptual_roughness * clamp01(m.index_of_refraction * 2.0f - 2.0f));
    f3 transmitted = sample_pyramid(fb, tu, tv, lod);
    // apply_volume_attenuation, lib.rs:275-290
    f3 attenuated = transmitted;
    if (p.attenuation_distance != __int_as_float(0x7f800000)) {
        float s = ray_length / p.attenuation_distance;
        attenuated.x *= expf(logf(p.attenuation_colour.x) * s);
        attenuated.y *= expf(logf(p.attenuation_colour.y) * s);
        attenuated.z *= expf(logf(p.attenuation_colour.z) * s);
    }
    float nov = dot3(p.normal, p.view);  // unclamped, lib.rs:345
    float2 brdf = sample_lut(lut, nov, m.perceptual_roughness);
    f3 f90 = add3(f0, df);
    f3 spec = add3(scale3(f0, brdf.x), scale3(f90, brdf.y));
    f3 one_m = mk3(1.0f - spec.x, 1.0f - spec.y, 1.0f - spec.z);
    return mul3(mul3(one_m, attenuated), m.diffuse_colour);
}

}  // namespace trd

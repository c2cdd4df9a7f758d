/* oracle/vecmath.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Scalar restatement of the glam 0.19.0 vector operations the reference's
 * per-pixel code relies on.  glam itself is NOT vendored in /root/reference
 * (Cargo.lock:601-607 pins glam 0.19.0 from crates.io), so the operation order
 * below restates glam's published scalar code path:
 *   dot3        = (ax*bx + ay*by) + az*bz
 *   normalize   = v * (1 / sqrt(dot(v,v)))
 *   lerp(a,b,s) = a + (b - a) * s
 *   Mat4 * Vec4 = ((c0*x + c1*y) + c2*z) + c3*w
 *   Quat * Vec3 = v*(w*w - b.b) + b*(2*(v.b)) + (b x v)*(2w),  b = q.xyz
 *   Vec3 / f32  = three divides
 * Build with -ffp-contract=off so no FMA contraction changes the rounding.
 * PARITY: pinned through its users to the reference's compiled shader modules (oracle.h).
 */
#ifndef ORACLE_VECMATH_H
#define ORACLE_VECMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y; } v2;
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;
typedef struct { v4 c[4]; } m4; /* column-major, like glam::Mat4 */

static inline v3 v3_new(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_splat(float s) { v3 r = {s, s, s}; return r; }
static inline v4 v4_new(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_new(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_new(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_new(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_scale(v3 a, float s) { return v3_new(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_divs(v3 a, float s) { return v3_new(a.x / s, a.y / s, a.z / s); }
static inline v3 v3_neg(v3 a) { return v3_new(-a.x, -a.y, -a.z); }
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float v3_length_squared(v3 a) { return v3_dot(a, a); }
static inline float v3_length(v3 a) { return sqrtf(v3_dot(a, a)); }
static inline v3 v3_normalize(v3 a) { return v3_scale(a, 1.0f / v3_length(a)); }
static inline v3 v3_lerp(v3 a, v3 b, float s) { return v3_add(a, v3_scale(v3_sub(b, a), s)); }
static inline float f_max(float a, float b) { return a > b ? a : (b != b ? a : b); } /* f32::max: NaN-ignoring */
static inline float f_min(float a, float b) { return a < b ? a : (b != b ? a : b); }
static inline float v3_max_element(v3 a) { return f_max(a.x, f_max(a.y, a.z)); }
static inline v3 v3_max(v3 a, v3 b) { return v3_new(f_max(a.x, b.x), f_max(a.y, b.y), f_max(a.z, b.z)); }
static inline v3 v3_min(v3 a, v3 b) { return v3_new(f_min(a.x, b.x), f_min(a.y, b.y), f_min(a.z, b.z)); }
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_new(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline v3 v3_one_minus(v3 a) { return v3_new(1.0f - a.x, 1.0f - a.y, 1.0f - a.z); }

static inline v4 m4_mul_v4(const m4* m, v4 v) {
    v4 r;
    r.x = ((m->c[0].x * v.x + m->c[1].x * v.y) + m->c[2].x * v.z) + m->c[3].x * v.w;
    r.y = ((m->c[0].y * v.x + m->c[1].y * v.y) + m->c[2].y * v.z) + m->c[3].y * v.w;
    r.z = ((m->c[0].z * v.x + m->c[1].z * v.y) + m->c[2].z * v.z) + m->c[3].z * v.w;
    r.w = ((m->c[0].w * v.x + m->c[1].w * v.y) + m->c[2].w * v.z) + m->c[3].w * v.w;
    return r;
}

static inline v3 quat_mul_v3(v4 q, v3 v) {
    v3 b = v3_new(q.x, q.y, q.z);
    float b2 = v3_dot(b, b);
    v3 r = v3_scale(v, q.w * q.w - b2);
    r = v3_add(r, v3_scale(b, v3_dot(v, b) * 2.0f));
    r = v3_add(r, v3_scale(v3_cross(b, v), q.w * 2.0f));
    return r;
}

/* IEEE binary16 <-> binary32, round-to-nearest-even (what an
 * R16G16B16A16_SFLOAT attachment store does; src/main.rs:2370). */
static inline uint16_t f32_to_f16_bits(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | (absx > 0x7f800000u ? 0x0200u | ((absx >> 13) & 0x3ffu) : 0u));
    }
    if (absx >= 0x477ff000u) { /* >= 65520 rounds to inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x38800000u) { /* subnormal half or zero */
        if (absx < 0x33000000u) return (uint16_t)sign; /* < 2^-25 -> 0 (2^-25 exactly ties to even = 0) */
        uint32_t e = absx >> 23;
        uint32_t m = (absx & 0x7fffffu) | 0x800000u;
        uint32_t shift = 126u - e; /* 14..24 */
        uint32_t half = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1u);
        if (rem > halfway || (rem == halfway && (half & 1u))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t r = absx + 0xc8000000u; /* rebias exponent: -(127-15)<<23 */
    uint32_t half = r >> 13;
    uint32_t rem = r & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) half++;
    return (uint16_t)(sign | half);
}

static inline float f16_bits_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) {
            x = sign;
        } else {
            int shift = 0;
            while (!(m & 0x400u)) { m <<= 1; shift++; }
            m &= 0x3ffu;
            x = sign | ((uint32_t)(113 - shift) << 23) | (m << 13);
        }
    } else if (e == 31) {
        x = sign | 0x7f800000u | (m << 13);
    } else {
        x = sign | ((e + 112u) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

#endif

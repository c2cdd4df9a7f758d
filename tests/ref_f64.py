"""Independent float64 numpy restatement of glam-pbr (second opinion for the C oracle).

Written from /root/reference/glam-pbr/src/lib.rs separately from oracle/pbr.c
and vectorised over a leading batch axis, so a transcription slip in one of the
two shows up as a disagreement far above fp32 rounding.
"""
import numpy as np

EPS = np.float64(np.finfo(np.float32).eps)


def dot(a, b):
    return np.sum(a * b, axis=-1)


def normalize(v):
    return v / np.sqrt(dot(v, v))[..., None]


def sdot(a, b):  # Dot::new, lib.rs:92-99
    return np.maximum(dot(a, b), EPS)


def d_ggx(noh, alpha):  # lib.rs:101-109
    a2 = alpha * alpha
    f = noh * noh * (a2 - 1.0) + 1.0
    return a2 / (np.pi * f * f)


def v_smith(nov, nol, alpha):  # lib.rs:114-133
    a2 = alpha * alpha
    gv = nol * np.sqrt(nov * nov * (1.0 - a2) + a2)
    gl = nov * np.sqrt(nol * nol * (1.0 - a2) + a2)
    g = gv + gl
    return np.where(g > 0.0, 0.5 / np.where(g > 0, g, 1.0), 0.0)


def fresnel(vdh, f0, f90):  # lib.rs:137-139
    return f0 + (f90 - f0) * ((1.0 - vdh) ** 5.0)[..., None]


def lerp(a, b, t):
    return a + (b - a) * t


def f0_f90(m):  # lib.rs:425-435
    root = (m["ior"] - 1.0) / (m["ior"] + 1.0)
    diel = (root * root)[..., None] * m["specular_colour"] * m["specular_factor"][..., None]
    f0 = lerp(diel, m["diffuse"], m["metallic"][..., None])
    f90 = lerp(np.repeat(m["specular_factor"][..., None], 3, -1), 1.0, m["metallic"][..., None])
    return f0, f90


def basic_brdf(n, l, li, v, m):  # lib.rs:377-423
    alpha = m["roughness"] ** 2
    h = normalize(v + l)
    noh, nov, nol, voh = sdot(n, h), sdot(n, v), sdot(n, l), sdot(v, h)
    c_diff = lerp(m["diffuse"], 0.0, m["metallic"][..., None])
    f0, f90 = f0_f90(m)
    F = fresnel(voh, f0, f90)
    diffuse = li * nol[..., None] * ((1.0 - F.max(axis=-1)) / np.pi)[..., None] * c_diff
    spec = li * nol[..., None] * (d_ggx(noh, alpha) * v_smith(nov, nol, alpha))[..., None] * F
    return diffuse, spec


def clamp01(x):
    return np.minimum(np.maximum(x, 0.0), 1.0)


def transmission_btdf(m, n, v, l):  # lib.rs:200-233
    alpha = m["roughness"] ** 2 * clamp01(m["ior"] * 2.0 - 2.0)
    lm = normalize(l + 2.0 * n * dot(-l, n)[..., None])
    h = normalize(v + lm)
    noh, voh, nov, nol = sdot(n, h), sdot(v, h), sdot(n, v), sdot(n, lm)
    f0, f90 = f0_f90(m)
    F = fresnel(voh, f0, f90)
    return (1.0 - F) * (d_ggx(noh, alpha) * v_smith(nov, nol, alpha))[..., None] * m["diffuse"]


def refract(i, n, ior):  # lib.rs:248-256
    eta = 1.0 / ior
    ndi = dot(n, i)
    k = 1.0 - eta * eta * (1.0 - ndi * ndi)
    return eta[..., None] * i - (eta * ndi + np.sqrt(k))[..., None] * n


def bilinear(img, u, v):
    """img (h,w,c) float64; clamp-to-edge, texel centres at +0.5."""
    h, w = img.shape[:2]
    px, py = u * w - 0.5, v * h - 0.5
    x0, y0 = np.floor(px), np.floor(py)
    fx, fy = (px - x0)[..., None], (py - y0)[..., None]
    x0i = np.clip(x0.astype(np.int64), 0, w - 1)
    x1i = np.clip(x0.astype(np.int64) + 1, 0, w - 1)
    y0i = np.clip(y0.astype(np.int64), 0, h - 1)
    y1i = np.clip(y0.astype(np.int64) + 1, 0, h - 1)
    top = lerp(img[y0i, x0i], img[y0i, x1i], fx)
    bot = lerp(img[y1i, x0i], img[y1i, x1i], fx)
    return lerp(top, bot, fy)


def sample_pyramid(levels, u, v, lod):
    """levels: list of float64 (h,w,4); scalar lod per batch allowed as array."""
    nl = len(levels)
    lod = np.clip(lod, 0.0, nl - 1.0)
    l0 = np.floor(lod).astype(np.int64)
    l1 = np.minimum(l0 + 1, nl - 1)
    t = (lod - l0)[..., None]
    out = np.zeros(u.shape + (3,))
    for lv in range(nl):
        m0, m1 = l0 == lv, l1 == lv
        if not (m0.any() or m1.any()):
            continue
        s = bilinear(levels[lv], u, v)[..., :3]
        out = out + np.where(m0[..., None], s * (1.0 - t), 0.0) + np.where(m1[..., None], s * t, 0.0)
    return out


def ibl_volume_refraction(m, size_x, n, v, pv, pos, thickness, scale, att_d, att_c, levels, lut):  # lib.rs:292-354
    r = normalize(refract(-v, n, m["ior"]))
    length = thickness * scale
    exit_p = pos + r * length[..., None]
    ph = np.concatenate([exit_p, np.ones(exit_p.shape[:-1] + (1,))], -1) @ pv.T
    uv = (ph[..., :2] / ph[..., 3:4] + 1.0) / 2.0
    lod = np.log2(float(size_x)) * m["roughness"] * clamp01(m["ior"] * 2.0 - 2.0)
    t = sample_pyramid(levels, uv[..., 0], uv[..., 1], lod)
    with np.errstate(divide="ignore", invalid="ignore"):
        coeff = -np.log(att_c) / att_d[..., None]
        trans = np.exp(-coeff * length[..., None])
    att = np.where(np.isinf(att_d)[..., None], t, trans * t)
    nov = dot(n, v)
    brdf = bilinear(lut, nov, m["roughness"])
    f0, f90 = f0_f90(m)
    spec = f0 * brdf[..., 0:1] + f90 * brdf[..., 1:2]
    return (1.0 - spec) * att * m["diffuse"]


def lottes(color, p):  # shader/src/tonemapping.rs:9-25
    mx = np.maximum(color.max(axis=-1), np.finfo(np.float32).tiny)

    def inner(x):
        z = x ** p["a"]
        return z / (z ** p["d"] * p["b"] + p["c"])

    ratio = color / mx[..., None]
    tm = inner(mx)
    ratio = ratio ** (p["saturation"] / p["cross_saturation"])
    ratio = lerp(ratio, 1.0, (tm ** p["crosstalk"])[..., None])
    ratio = ratio ** p["cross_saturation"]
    return np.clip(ratio * tm[..., None], 0.0, 1.0)

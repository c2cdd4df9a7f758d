"""Property tests of the CPU oracle (hypothesis): invariants that follow from the reference's formulas and hold for any
input, used to pin the restatement beyond the known-answer points (the reference ships no tests)."""
import math

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from transmission_renderer_b200 import abi, host, scenes

f32 = np.float32
unit = st.floats(-1.0, 1.0, allow_nan=False)


def _vec(x, y, z, fallback):
    v = np.array([x, y, z], np.float64)
    n = np.linalg.norm(v)
    return (v / n if n > 1e-3 else np.array(fallback, np.float64)).astype(f32)


def _params(n, l, v, rough, metallic, base, ior=1.5, intensity=(1.0, 1.0, 1.0)):
    p = np.zeros(1, dtype=abi.basic_brdf_params)
    p["normal"], p["light"], p["view"] = n, l, v
    p["light_intensity"] = intensity
    m = p["material_params"]
    m["diffuse_colour"] = base
    m["metallic"] = metallic
    m["perceptual_roughness"] = rough
    m["index_of_refraction"] = ior
    m["specular_colour"] = (1, 1, 1)
    m["specular_factor"] = 1.0
    return p


@settings(max_examples=300, deadline=None)
@given(unit, unit, unit, unit, unit, unit, st.floats(0.08, 1.0), st.sampled_from([0.0, 1.0]),
       st.floats(0.05, 1.0))
def test_brdf_is_finite_non_negative_and_linear_in_the_light(oracle, lx, ly, lz, vx, vy, vz, rough, metallic, base):
    n = np.array([0, 0, 1], f32)
    l, v = _vec(lx, ly, abs(lz) + 0.05, (0, 0, 1)), _vec(vx, vy, abs(vz) + 0.05, (0, 0, 1))
    r1 = oracle.eval_basic_brdf(_params(n, l, v, rough, metallic, (base, base, base)))
    r3 = oracle.eval_basic_brdf(_params(n, l, v, rough, metallic, (base, base, base), intensity=(3.0, 3.0, 3.0)))
    for k in ("diffuse", "specular"):
        assert np.isfinite(r1[k]).all() and (r1[k] >= 0).all()
        np.testing.assert_allclose(r3[k], 3.0 * r1[k], rtol=2e-6, atol=1e-30)   # basic_brdf scales with light_intensity
    # Lambert lobe never exceeds base / pi * n.l (diffuse_brdf, glam-pbr lib.rs:356-360: it is scaled by 1 - max(F) <= 1)
    assert (r1["diffuse"] <= base / math.pi * max(float(l[2]), 1.2e-7) * (1 + 1e-5) + 1e-12).all()
    if metallic == 1.0:
        assert (r1["diffuse"] == 0).all()                                        # c_diff = lerp(base, 0, metallic)


@settings(max_examples=200, deadline=None)
@given(unit, unit, unit, unit, unit, unit, st.floats(0.1, 1.0))
def test_specular_lobe_is_reciprocal(oracle, lx, ly, lz, vx, vy, vz, rough):
    """D, V and F of the GGX lobe are symmetric in (l, v): specular / n.l with the roles swapped agrees (f_r(l,v) = f_r(v,l))."""
    n = np.array([0, 0, 1], f32)
    l, v = _vec(lx, ly, abs(lz) + 0.1, (0, 0, 1)), _vec(vx, vy, abs(vz) + 0.1, (0.6, 0, 0.8))
    a = oracle.eval_basic_brdf(_params(n, l, v, rough, 0.0, (0.5, 0.5, 0.5)))["specular"][0] / l[2]
    b = oracle.eval_basic_brdf(_params(n, v, l, rough, 0.0, (0.5, 0.5, 0.5)))["specular"][0] / v[2]
    np.testing.assert_allclose(a, b, rtol=5e-4)


@settings(max_examples=200, deadline=None)
@given(unit, unit, unit, st.floats(0.1, 1.0), st.floats(1.0, 2.5))
def test_btdf_of_a_light_behind_equals_the_mirrored_reflection(oracle, vx, vy, vz, rough, ior):
    """transmission_btdf mirrors the light about the surface (glam-pbr lib.rs:211): a light at -l behaves like the
    reflection lobe of l with alpha scaled by clamp(2 ior - 2, 0, 1) — and the result never depends on the sign trick."""
    n = np.array([0, 0, 1], f32)
    v = _vec(vx, vy, abs(vz) + 0.1, (0, 0, 1))
    l = _vec(0.3, -0.2, 0.9, (0, 0, 1))
    p = np.zeros(2, dtype=abi.transmission_btdf_params)
    for k in range(2):
        m = p["material_params"][k]
        m["diffuse_colour"] = (0.7, 0.8, 0.9)
        m["metallic"] = 0.0
        m["perceptual_roughness"] = rough
        m["index_of_refraction"] = ior
        m["specular_colour"] = (1, 1, 1)
        m["specular_factor"] = 1.0
    p["normal"], p["view"] = n, v
    p["light"][0] = l * np.array([1, 1, -1], f32)     # behind the surface
    p["light"][1] = l * np.array([1, 1, -1], f32)
    r = oracle.eval_transmission_btdf(p)
    assert np.isfinite(r).all() and (r >= 0).all()
    np.testing.assert_array_equal(r[0], r[1])
    assert (r[0] <= np.array([0.7, 0.8, 0.9], f32) * 1e6).all()


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40), st.integers(1, 40), st.integers(0, 255))
def test_mip_chain_of_a_constant_image_is_constant(oracle, w, h, v):
    img = np.full((h, w, 4), f32(v / 16.0), f32)
    levels = oracle.build_pyramid(oracle.f16_bits(img))
    assert len(levels) == host.mip_levels_for_size(w, h)
    for lv in levels:
        assert (lv == levels[0][0, 0]).all()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_mip_level_of_even_sizes_is_the_box_mean(oracle, k, seed):
    rng = np.random.default_rng(seed)
    w, h = 2 * k, 2 * (7 - k)
    img = rng.uniform(0, 8, (h, w, 4)).astype(np.float16).astype(f32)
    lv = oracle.build_pyramid(oracle.f16_bits(img))
    mean = img.reshape(h // 2, 2, w // 2, 2, 4).astype(np.float64).mean(axis=(1, 3))
    got = oracle.f16_to_f32(lv[1]).astype(np.float64)
    assert np.abs(got - mean).max() <= np.abs(mean).max() * 2.0 ** -10


@settings(max_examples=100, deadline=None)
@given(st.floats(-40, 40), st.floats(-10, 30), st.floats(-40, 40), st.floats(0.1, 3.0))
def test_cull_is_monotone_in_the_radius(oracle, x, y, z, scale):
    """A sphere that is visible stays visible when it grows (every test of `cull` compares against +radius,
    shader/src/lib.rs:455-463), and a sphere around the camera is always visible."""
    cam = scenes.Camera(640, 360, (0.0, 6.0, 0.0), 20.0, -10.0)
    prims = np.zeros(1, dtype=abi.primitive_info)
    prims["packed_bounding_sphere"] = (0, 0, 0, 1.0)
    prims["index_count"] = 3
    inst = np.zeros(3, dtype=abi.instance)
    inst["rotation"] = (0, 0, 0, 1)
    inst["translation_and_scale"][0] = (x, y, z, scale)
    inst["translation_and_scale"][1] = (x, y, z, scale * 2.0)
    inst["translation_and_scale"][2] = (0.0, 6.0, 0.0, 0.5)
    _, visible = oracle.frustum_culling(inst, prims, cam.culling())
    vis = set(int(i) for i in visible)
    assert 2 in vis
    assert (0 not in vis) or (1 in vis)


@settings(max_examples=100, deadline=None)
@given(st.floats(0.0, 60000.0), st.floats(0.0, 1.0), st.floats(0.0, 1.0))
def test_tonemap_is_bounded_and_monotone_in_exposure(oracle, peak, g, b):
    params = host.default_tonemap_params()
    hdr = np.zeros((1, 2, 4), f32)
    hdr[0, 0, :3] = (peak, peak * g, peak * b)
    hdr[0, 1, :3] = (2 * peak, 2 * peak * g, 2 * peak * b)
    hdr[..., 3] = 1.0
    out = oracle.tonemap_frame(oracle.f16_bits(np.minimum(hdr, 65504.0)), params).astype(int)
    assert out.min() >= 0 and out.max() <= 255 and (out[..., 3] == 255).all()
    assert out[0, 1, 0] >= out[0, 0, 0] - 1      # brighter input never comes out darker in its peak channel

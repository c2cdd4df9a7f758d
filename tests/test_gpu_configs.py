"""BASELINE.json's configurations at their STATED sizes, through libtr.so, against the oracle.

config 1 (512^2 synthetic G-buffer)            -> tests/test_gpu_parity.py::test_config1_* and test_gpu_reference_spirv.py
config 2 (1080p, 64 UV-spheres, 4 point lights, opaque only, K1-K4)          here, whole frame
config 3 (1080p, displaced torus knot over the sphere scene, K1-K6)          here, whole frame
config 4 (3840x2160, 10 k instances, 64 lights)                              here: oracle on three bands (top, middle, bottom)
config 5 (7680x4320, one of the 64 orbit views)                              here: oracle on one band
plus the multi-GPU determinism requirement on ONE device: the 4K frame rendered as 2, 4 and 8 bands —
each band through exactly the calls a rank makes — must equal the whole-frame render byte for byte.
"""
import numpy as np
import pytest

from pipeline import REL_L2_TOL, SRGB_TOL, gpu_setup, oracle_cluster_lights, oracle_scene, rel_l2
from transmission_renderer_b200 import Renderer, abi, host, scenes

pytestmark = pytest.mark.gpu


def _upload(r, lut, s):
    gpu_setup(r, lut, s["uniforms"], s["materials"], s["lights"])
    r.set_instances(s["instances"])
    r.set_primitives(s["primitives"])
    m = s["mesh"]
    r.set_mesh(m["positions"], m["normals"], m["uvs"], m["indices"])
    r.build_clusters(s["camera"].write_cluster_data())


def _oracle_bands(oracle, lut, s, bands, gpu_opaque16):
    """The oracle chain on rows `bands` of the frame.  The mip pyramid needs the whole opaque frame: rows outside the bands
    are taken from the GPU's opaque target (they only enter through the refraction fetch), the bands' own rows are the
    oracle's."""
    cam = s["camera"]
    pc = cam.push_constants()
    _, visible = oracle.frustum_culling(s["instances"], s["primitives"], cam.culling())
    _, cc, ci = oracle_cluster_lights(oracle, cam, s["uniforms"], s["lights"])
    sc = oracle_scene(pc, s["uniforms"], s["materials"], s["lights"], cc, ci)
    opaque16 = np.array(gpu_opaque16, copy=True)
    per_band = []
    for (y0, y1) in bands:
        g0, g1 = oracle.visibility(s["mesh"], s["instances"], s["primitives"], visible, pc, y0, y1)
        o32, o16 = oracle.shade_opaque_frame(g0, sc, y0, y1)
        opaque16[y0:y1] = o16[y0:y1]
        per_band.append((g0, g1, o32, o16))
    levels = oracle.build_pyramid(opaque16)
    out = []
    for (y0, y1), (g0, g1, o32, o16) in zip(bands, per_band):
        t32, t16 = oracle.shade_transmission_frame(g1, sc, levels, lut, o32, o16, y0, y1)
        out.append(dict(g0=g0, g1=g1, o32=o32, t32=t32, t16=t16))
    return out


def _check_bands(oracle, lut, s, bands, label):
    cam = s["camera"]
    w, h = cam.width, cam.height
    params = host.default_tonemap_params()
    with Renderer(w, h, f32_debug=True) as r:
        _upload(r, lut, s)
        r.frame(cam.frame_params(params))
        got32, srgb = r.read_hdr_f32(), r.read_srgb8()
        gg = [r.read_gbuffer(0), r.read_gbuffer(1)]
        opaque16 = r.read_pyramid_level(0)
    ref = _oracle_bands(oracle, lut, s, bands, opaque16)
    for (y0, y1), b in zip(bands, ref):
        for layer, key in ((0, "g0"), (1, "g1")):
            for plane in ("depth", "normal", "uv", "material_id"):
                a = np.asarray(gg[layer][plane]).reshape(h, w, -1)[y0:y1]
                o = np.asarray(b[key][plane]).reshape(h, w, -1)[y0:y1]
                assert a.tobytes() == o.tobytes(), f"{label} rows [{y0},{y1}) layer {layer} plane {plane}: {(a != o).sum()} values differ"
        e = rel_l2(got32[y0:y1, :, :3], b["t32"][y0:y1, :, :3])
        ref_srgb = oracle.tonemap_frame(b["t16"], params, y0, y1)
        d = np.abs(srgb[y0:y1].astype(int) - ref_srgb[y0:y1].astype(int))
        covered = float((np.asarray(b["g0"]["depth"]).reshape(h, w)[y0:y1] != 0).mean())
        print(f"{label} rows [{y0},{y1}): G-buffer bit-exact, opaque coverage {covered:.2f}, final fp32 rel-L2 {e:.2e}, sRGB8 max diff {d.max()}")
        assert e < REL_L2_TOL
        assert d.max() <= SRGB_TOL


def test_config2_1080p_spheres_whole_frame(oracle, ggx_lut):
    """configs[1]: opaque only — cull, light lists, visibility, `fragment` over the whole 1920x1080 frame."""
    s = scenes.sphere_grid_scene(1920, 1080, transmissive_knot=False)
    assert len(s["lights"]) == 4
    _check_bands(oracle, ggx_lut, s, [(0, 1080)], "config 2")


def test_config3_1080p_knot_whole_frame(oracle, ggx_lut):
    """configs[2]: the displaced torus knot (512 x 64 segments, roughness 0.25, volume attenuation) over the sphere scene:
    opaque pass -> mip chain -> transmission pass -> tonemap, whole 1920x1080 frame."""
    s = scenes.sphere_grid_scene(1920, 1080, transmissive_knot=True)
    _check_bands(oracle, ggx_lut, s, [(0, 1080)], "config 3")


def test_config4_4k_three_bands(oracle, ggx_lut):
    """configs[3] at 3840x2160 / 10 000 instances / 64 lights: the oracle on a top, a middle and a bottom band."""
    s = scenes.instanced_scene(3840, 2160)
    _check_bands(oracle, ggx_lut, s, [(0, 40), (1060, 1100), (2120, 2160)], "config 4")


def test_config5_8k_one_view(oracle, ggx_lut):
    """configs[4]: one of the 64 orbit views at 7680x4320 (13 mip levels); the oracle on one band."""
    s = scenes.instanced_scene(7680, 4320, yaw_deg=360.0 * 5 / 64)
    _check_bands(oracle, ggx_lut, s, [(2000, 2024)], "config 5 view 5")


@pytest.mark.parametrize("n_bands", [2, 4, 8])
def test_4k_bands_equal_whole_frame_bitwise(ggx_lut, n_bands):
    """SURVEY.md 8e determinism: N-GPU output == 1-GPU output bitwise.  Emulated on one device: every band goes through the
    calls a rank makes (cull, light lists, band visibility, band opaque shading; then the shared pyramid; band transmission,
    band tonemap), and G-buffer, HDR (RGBA16F) and sRGB8 bytes must equal the whole-frame render."""
    w, h = 3840, 2160
    s = scenes.instanced_scene(w, h)
    cam = s["camera"]
    pc, params = cam.push_constants(), host.default_tonemap_params()
    with Renderer(w, h) as r:
        _upload(r, ggx_lut, s)
        r.frame(cam.frame_params(params))
        whole = dict(hdr=r.read_hdr(), srgb=r.read_srgb8(), g0=r.read_gbuffer(0), g1=r.read_gbuffer(1), mip0=r.read_pyramid_level(0))
    bands = [host.band_rows(h, b, n_bands) for b in range(n_bands)]
    with Renderer(w, h) as r:
        _upload(r, ggx_lut, s)
        srgb = np.zeros((h, w, 4), np.uint8)
        for (y0, y1) in bands:
            r.set_band(y0, y1)
            r.cull(cam.culling())
            r.assign_lights(cam.assign_lights())
            r.visibility(pc)
            r.shade_opaque(pc)
        r.generate_mips()
        for (y0, y1) in bands:
            r.set_band(y0, y1)
            r.shade_transmission(pc)
            r.tonemap(params)
            r.read_srgb8(srgb)
        hdr = r.read_hdr()
        mip0 = r.read_pyramid_level(0)
    assert mip0.tobytes() == whole["mip0"].tobytes(), "opaque frame differs"
    assert hdr.tobytes() == whole["hdr"].tobytes(), f"HDR differs in {(hdr != whole['hdr']).any(axis=2).sum()} pixels"
    assert srgb.tobytes() == whole["srgb"].tobytes()

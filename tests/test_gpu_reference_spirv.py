"""The CUDA path against outputs of the reference's OWN compiled shaders.

tests/golden/spirv_golden.npz holds what the reference's shipped SPIR-V modules produce on the cases of
tests/spirv_cases.py (executed on the CPU through oracle/spv2c.py, see tests/golden/make_spirv_golden.py and
tests/test_reference_spirv.py).  Here the same cases go through libtr.so on the GPU: cull / draw lists / cluster AABBs /
light lists bit-exact, shaded fp32 pixels within the north-star tolerance of the modules' pixels.  The fixture is all
this file needs — neither /root/reference nor oracle/_ref exists on the GPU box's path of this test.
"""
import numpy as np
import pytest

import spirv_cases as cases
from pipeline import REL_L2_TOL, gpu_setup, rel_l2
from transmission_renderer_b200 import Renderer, abi, host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return np.load(cases.GOLDEN)


def _upload(r, lut, s):
    gpu_setup(r, lut, s["uniforms"], s["materials"], s["lights"])
    r.set_instances(s["instances"])
    r.set_primitives(s["primitives"])
    m = s["mesh"]
    r.set_mesh(m["positions"], m["normals"], m["uvs"], m["indices"])
    r.build_clusters(s["camera"].write_cluster_data())


def test_compute_stages_equal_the_reference_modules(golden, ggx_lut):
    """K1 / K2 against frustum_culling.spv, demultiplex_draws.spv, write_cluster_data.spv, assign_lights_to_clusters.spv
    on BASELINE config 4's scene (10 k instances, 64 point lights + 6 spotlights)."""
    s = cases.instanced()
    cam = s["camera"]
    with Renderer(cam.width, cam.height) as r:
        _upload(r, ggx_lut, s)
        r.cull(cam.culling())
        r.assign_lights(cam.assign_lights())
        visible = r.read_visible_instances()
        counts = r.read_instance_counts(len(s["primitives"]))
        draws = [r.read_draws(b) for b in range(4)]
        n_clusters = len(golden["instanced_light_counts"])
        aabbs = r.read_cluster_aabbs(n_clusters)
        lc, li = r.read_cluster_lights(n_clusters)
    assert np.array_equal(visible, golden["instanced_visible"])
    assert np.array_equal(counts, golden["instanced_instance_counts"])
    assert np.array_equal(np.array([len(d) for d in draws], np.uint32), golden["instanced_draw_counts"])
    assert np.array_equal(np.concatenate([np.frombuffer(d.tobytes(), np.uint32) for d in draws]), golden["instanced_draws"])
    assert np.array_equal(np.frombuffer(aabbs.tobytes(), np.uint32), golden["instanced_aabbs"])
    assert np.array_equal(lc, golden["instanced_light_counts"])
    lists = np.concatenate([li[c * 128:c * 128 + n] for c, n in enumerate(lc)])
    assert np.array_equal(lists, golden["instanced_light_lists"])


@pytest.mark.parametrize("name,make", [("instanced", cases.instanced), ("spheres", cases.spheres)])
def test_frames_match_the_reference_modules(golden, ggx_lut, name, make):
    """K1-K6 through tr_frame; the shaded fp32 pixels against fragment.spv / fragment_transmission.spv on the sampled
    pixels: rel-L2 <= 1e-4 (north star) and no element off by more than 1e-3 of the frame's scale."""
    s = make()
    cam = s["camera"]
    with Renderer(cam.width, cam.height, f32_debug=True) as r:
        _upload(r, ggx_lut, s)
        r.frame(cam.frame_params(host.default_tonemap_params(), flags=abi.TR_FRAME_SKIP_TONEMAP))
        g0, g1 = r.read_gbuffer(0), r.read_gbuffer(1)
        final = r.read_hdr_f32().reshape(-1, 4)
        r.cull(cam.culling())
        r.assign_lights(cam.assign_lights())
        r.visibility(cam.push_constants())
        r.shade_opaque(cam.push_constants())
        opaque = r.read_hdr_f32().reshape(-1, 4)
    assert np.array_equal(cases.sample_pixels(g0["depth"]), golden[f"{name}_opaque_px"])          # same coverage as the fixture's G-buffer
    assert np.array_equal(cases.sample_pixels(g1["depth"]), golden[f"{name}_transmission_px"])
    for label, got, px, ref in (("opaque", opaque, golden[f"{name}_opaque_px"], golden[f"{name}_opaque_rgba"]),
                                ("transmission", final, golden[f"{name}_transmission_px"], golden[f"{name}_transmission_rgba"])):
        e = rel_l2(got[px][:, :3], ref[:, :3])
        scale = float(np.sqrt(np.mean(ref[:, :3].astype(np.float64) ** 2)))
        worst = float(np.abs(got[px][:, :3].astype(np.float64) - ref[:, :3]).max() / scale)
        print(f"{name} {label}: rel-L2 vs the reference's module {e:.2e}, worst element {worst:.2e} of the frame's rms")
        assert e < REL_L2_TOL and worst < 1e-3


def test_config1_matches_the_reference_module(golden, ggx_lut):
    """BASELINE config 1 (synthetic G-buffer, roughness 0.25, sun + one light) against fragment_transmission.spv."""
    from oracle import pyoracle as po
    s = cases.config1()
    cam = s["camera"]
    size = s["gbuffer"]["depth"].shape[0]
    with Renderer(size, size, f32_debug=True) as r:
        gpu_setup(r, ggx_lut, s["uniforms"], s["materials"], s["lights"])
        r.build_clusters(cam.write_cluster_data())
        r.assign_lights(cam.assign_lights())
        r.set_gbuffer(abi.TR_LAYER_TRANSMISSIVE, s["gbuffer"])
        r.set_opaque_frame(po.f16_bits(s["opaque"]))
        r.generate_mips()
        r.shade_transmission(cam.push_constants())
        got = r.read_hdr_f32().reshape(-1, 4)
    px, ref = golden["config1_transmission_px"], golden["config1_transmission_rgba"]
    e = rel_l2(got[px][:, :3], ref[:, :3])
    print(f"config 1: rel-L2 vs the reference's module {e:.2e}")
    assert e < REL_L2_TOL

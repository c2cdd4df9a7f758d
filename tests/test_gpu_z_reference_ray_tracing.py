"""N4 on the GPU against the reference's ray-tracing builds of the fragment modules (compiled-shaders/ray-tracing/*.spv).

tests/golden/spirv_golden.npz (keys shadows_*) holds the pixels those modules produce on tests/spirv_cases.shadows() — executed
on the CPU through oracle/spv2c.py with the shadow-ray definition of oracle/shadow.c as their ray-query environment
(tests/golden/make_spirv_golden.py, tests/test_reference_spirv.py::test_golden_ray_tracing_fragments).  Only the fixture is
needed here."""
import numpy as np
import pytest

import spirv_cases as cases
from pipeline import REL_L2_TOL, gpu_setup, rel_l2
from transmission_renderer_b200 import Renderer, host

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


def test_shadowed_frame_matches_the_reference_ray_tracing_modules(golden, ggx_lut):
    """N4: one frame with ray-queried shadows (the sequence of __graft_entry__.smoke) against the pixels the reference's
    ray-tracing builds of `fragment` / `fragment_transmission` produced (compiled-shaders/ray-tracing/*.spv, executed on the
    CPU with the shadow-ray definition of oracle/shadow.c as their ray-query environment)."""
    s = cases.shadows()
    cam = s["camera"]
    with Renderer(cam.width, cam.height, f32_debug=True) as r:
        _upload(r, ggx_lut, s)
        handle = r.build_acceleration_structures()
        r.frame(cam.frame_params(host.default_tonemap_params(), acceleration_structure_address=handle))
        g0, g1 = r.read_gbuffer(0), r.read_gbuffer(1)
        final = r.read_hdr_f32().reshape(-1, 4)
    px_o, px_t = golden["shadows_opaque_px"], golden["shadows_transmission_px"]
    assert np.array_equal(cases.sample_pixels(g0["depth"]), px_o) and np.array_equal(cases.sample_pixels(g1["depth"]), px_t)
    # the final frame holds the opaque result wherever no glass covers it, the transmissive result where it does
    bare = np.asarray(g1["depth"]).reshape(-1)[px_o] == 0
    assert bare.sum() > 500 and len(px_t) > 30
    e_o = rel_l2(final[px_o[bare]][:, :3], golden["shadows_opaque_rgba"][bare][:, :3])
    e_t = rel_l2(final[px_t][:, :3], golden["shadows_transmission_rgba"][:, :3])
    print(f"shadows: rel-L2 vs the reference's ray-tracing modules, opaque {e_o:.2e}, transmission {e_t:.2e}")
    # measured on B200: opaque 1.7e-7, transmission 8.3e-7
    assert e_o < REL_L2_TOL and e_t < REL_L2_TOL

"""The cases whose outputs are pinned by the reference's shipped SPIR-V (tests/golden/spirv_golden.npz).

One definition used three times: by tests/golden/make_spirv_golden.py (which EXECUTES the reference's modules through
oracle/spvref.py and stores their outputs), by tests/test_reference_spirv.py (C oracle vs those outputs, and vs the live
modules where oracle/_ref exists) and by the -m gpu tests (CUDA vs the same outputs)."""
import os

import numpy as np

from transmission_renderer_b200 import host, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "spirv_golden.npz")
SAMPLE_STRIDE = 13  # every 13th covered pixel of a frame is stored


def lut_rgba8():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ggx_lut_rg.npz"))
    lut = np.zeros(z["rg"].shape[:2] + (4,), np.uint8)
    lut[..., :2] = z["rg"]
    lut[..., 3] = 255
    return lut


def instanced(width=480, height=270, spotlights=True):
    """BASELINE config 4's scene (10 k instances, 64 lights) at a test size, plus 6 spotlights so that the cone culling of
    assign_lights_to_clusters and the spotlight factor of `fragment` are on the pinned path."""
    s = scenes.instanced_scene(width, height)
    if spotlights:
        spots = [host.light_new_spot((x, 9.0, z), c, 40.0, scenes._unit(d), 0.3, 0.5)
                 for x, z, c, d in [(-8, -14, (1, .9, .8), (0.2, -1, -0.3)), (6, -18, (.8, .9, 1), (-0.1, -1, 0.2)),
                                    (0, -9, (1, 1, 1), (0, -1, -0.5)), (14, -25, (1, .7, .7), (-0.4, -1, 0)),
                                    (-15, -22, (.7, 1, .7), (0.3, -1, 0.1)), (3, -30, (.7, .7, 1), (0, -1, 0.4))]]
        s["lights"] = np.concatenate([s["lights"]] + spots)
    return s


def spheres(width=320, height=180):
    """BASELINE config 2/3's scene: 64 UV-spheres, ground, the four point lights, the transmissive torus knot."""
    return scenes.sphere_grid_scene(width, height, transmissive_knot=True)


def config1(size=192):
    return scenes.config1(size, 0.25, lights=[host.light_new_point((0.5, 3.0, 1.5), (1.0, 0.8, 0.6), 8.0)])


def shadows(width=192, height=108):
    """N4: occluders over a ground quad, sun + point lights + a spotlight, an alpha-clip caster and a glass receiver."""
    return scenes.shadow_scene(width, height)


def sample_pixels(depth):
    """Flat indices of every SAMPLE_STRIDE-th covered pixel."""
    return np.nonzero(np.asarray(depth).reshape(-1) != 0)[0][::SAMPLE_STRIDE]


def run_compute(impl, s):
    """cull -> demultiplex -> cluster AABBs -> light lists through `impl` (pyoracle or spvref)."""
    cam = s["camera"]
    if hasattr(impl, "frustum_culling_visible"):
        counts = impl.frustum_culling(s["instances"], s["primitives"], cam.culling())
        visible = impl.frustum_culling_visible(s["instances"], s["primitives"], cam.culling())
    else:
        counts, visible = impl.frustum_culling(s["instances"], s["primitives"], cam.culling())
    draws, draw_counts = impl.demultiplex_draws(s["primitives"], counts)
    aabbs = impl.write_cluster_data(s["uniforms"], cam.write_cluster_data())
    light_counts, light_indices = impl.assign_lights_to_clusters(s["lights"], aabbs, cam.assign_lights())
    return dict(instance_counts=counts, visible=visible, draw_counts=draw_counts,
                draws=np.concatenate([np.frombuffer(d.tobytes(), np.uint32) for d in draws]),
                aabbs=np.frombuffer(aabbs.tobytes(), np.uint32).copy(), light_counts=light_counts,
                light_lists=np.concatenate([light_indices[c * 128:c * 128 + n] for c, n in enumerate(light_counts)]
                                           + [np.zeros(0, np.uint32)]))


def shade_scene(s, light_counts, light_indices):
    pc = s["camera"].push_constants()
    return dict(push_constants=pc, uniforms=s["uniforms"], materials=s["materials"], lights=s["lights"],
                cluster_light_counts=light_counts, cluster_light_indices=light_indices, textures=s.get("textures"))

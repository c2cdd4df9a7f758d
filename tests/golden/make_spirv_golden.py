"""Writes tests/golden/spirv_golden.npz by EXECUTING the reference's shipped SPIR-V modules
(/root/reference/compiled-shaders/{normal,ray-tracing}/*.spv -> oracle/spv2c.py -> oracle/_ref/libspvref.so) on the cases of
tests/spirv_cases.py.  These are outputs of the reference's own compiled code; the C oracle and the CUDA path are checked
against them.  G-buffers (rasterisation) and sampled images are inputs here — fixed-function in the reference, defined by
SURVEY.md Appendix E.  Run in the build container:  python tests/golden/make_spirv_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spirv_cases as cases  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from oracle import spvref  # noqa: E402

out = {}
lut = cases.lut_rgba8()

# --- config 4 scene: compute stages + both fragment stages
s = cases.instanced()
r = cases.run_compute(spvref, s)
for k, v in r.items():
    out["instanced_" + k] = v
counts, indices = spvref.assign_lights_to_clusters(s["lights"], spvref.write_cluster_data(s["uniforms"], s["camera"].write_cluster_data()),
                                                   s["camera"].assign_lights())
g0, g1 = po.visibility(s["mesh"], s["instances"], s["primitives"], r["visible"], s["camera"].push_constants())
sc = cases.shade_scene(s, counts, indices)
o32, o16 = spvref.shade_opaque_frame(g0, sc)
t32, _ = spvref.shade_transmission_frame(g1, sc, po.build_pyramid(o16), lut, o32, o16)
p0, p1 = cases.sample_pixels(g0["depth"]), cases.sample_pixels(g1["depth"])
out["instanced_opaque_px"], out["instanced_opaque_rgba"] = p0.astype(np.uint32), o32.reshape(-1, 4)[p0]
out["instanced_transmission_px"], out["instanced_transmission_rgba"] = p1.astype(np.uint32), t32.reshape(-1, 4)[p1]

# --- config 2/3 scene
s = cases.spheres()
r = cases.run_compute(spvref, s)
for k in ("visible", "light_counts", "light_lists"):
    out["spheres_" + k] = r[k]
counts, indices = spvref.assign_lights_to_clusters(s["lights"], spvref.write_cluster_data(s["uniforms"], s["camera"].write_cluster_data()),
                                                   s["camera"].assign_lights())
g0, g1 = po.visibility(s["mesh"], s["instances"], s["primitives"], r["visible"], s["camera"].push_constants())
sc = cases.shade_scene(s, counts, indices)
o32, o16 = spvref.shade_opaque_frame(g0, sc)
t32, _ = spvref.shade_transmission_frame(g1, sc, po.build_pyramid(o16), lut, o32, o16)
p0, p1 = cases.sample_pixels(g0["depth"]), cases.sample_pixels(g1["depth"])
out["spheres_opaque_px"], out["spheres_opaque_rgba"] = p0.astype(np.uint32), o32.reshape(-1, 4)[p0]
out["spheres_transmission_px"], out["spheres_transmission_rgba"] = p1.astype(np.uint32), t32.reshape(-1, 4)[p1]

# --- config 1: synthetic G-buffer, sun + one point light, roughness 0.25
s = cases.config1()
cam = s["camera"]
aabbs = spvref.write_cluster_data(s["uniforms"], cam.write_cluster_data())
counts, indices = spvref.assign_lights_to_clusters(s["lights"], aabbs, cam.assign_lights())
sc = cases.shade_scene(s, counts, indices)
size = s["gbuffer"]["depth"].shape[0]
levels = po.build_pyramid(po.f16_bits(s["opaque"]))
t32, _ = spvref.shade_transmission_frame(s["gbuffer"], sc, levels, lut, np.zeros((size, size, 4), np.float32),
                                         np.zeros((size, size, 4), np.uint16))
p1 = cases.sample_pixels(s["gbuffer"]["depth"])
out["config1_transmission_px"], out["config1_transmission_rgba"] = p1.astype(np.uint32), t32.reshape(-1, 4)[p1]

# --- N4: the ray-tracing builds of both fragment modules (compiled-shaders/ray-tracing/*.spv) on the shadow scene; the
# acceleration structure and the ray / triangle rule are the environment's (oracle/shadow.c), everything else the modules'
s = cases.shadows()
cam = s["camera"]
_, visible = po.frustum_culling(s["instances"], s["primitives"], cam.culling())
counts, indices = spvref.assign_lights_to_clusters(s["lights"], spvref.write_cluster_data(s["uniforms"], cam.write_cluster_data()),
                                                   cam.assign_lights())
g0, g1 = po.visibility(s["mesh"], s["instances"], s["primitives"], visible, cam.push_constants())
sc = cases.shade_scene(s, counts, indices)
sc["accel"] = po.Accel(s["mesh"], s["instances"], s["primitives"])
o32, o16 = spvref.shade_opaque_frame(g0, sc)
t32, _ = spvref.shade_transmission_frame(g1, sc, po.build_pyramid(o16), lut, o32, o16)
p0, p1 = cases.sample_pixels(g0["depth"]), cases.sample_pixels(g1["depth"])
out["shadows_opaque_px"], out["shadows_opaque_rgba"] = p0.astype(np.uint32), o32.reshape(-1, 4)[p0]
out["shadows_transmission_px"], out["shadows_transmission_rgba"] = p1.astype(np.uint32), t32.reshape(-1, 4)[p1]

np.savez_compressed(cases.GOLDEN, **out)
print({k: v.shape for k, v in out.items()}, os.path.getsize(cases.GOLDEN), "bytes")

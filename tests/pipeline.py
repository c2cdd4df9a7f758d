"""Shared helpers of the parity tests: run the same inputs through the CPU oracle and through
libtr.so, and the comparison metrics (north_star: relative L2 <= 1e-4 on linear HDR, <= 2/255 per
sRGB8 channel, bit-exact for every integer / index result)."""
import numpy as np

from transmission_renderer_b200 import abi, host

REL_L2_TOL = 1e-4      # BASELINE.json north_star
SRGB_TOL = 2           # /255 per channel


def rel_l2(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    ok = np.isfinite(ref) & np.isfinite(got)
    # non-finite values (fp16 overflow at singular silhouette pixels) must coincide
    bad = np.isfinite(ref) != np.isfinite(got)
    assert bad.mean() < 1e-5, f"non-finite mismatch on {bad.sum()} values"
    diff = got[ok] - ref[ok]
    den = np.linalg.norm(ref[ok])
    return float(np.linalg.norm(diff) / den) if den > 0 else float(np.linalg.norm(diff))


def oracle_scene(pc, uniforms, materials, lights, counts, indices):
    return dict(push_constants=pc, uniforms=uniforms, materials=materials, lights=lights,
                cluster_light_counts=counts, cluster_light_indices=indices)


def oracle_cluster_lights(oracle, cam, uniforms, lights):
    aabbs = oracle.write_cluster_data(uniforms, cam.write_cluster_data())
    if len(lights) == 0:
        n = len(aabbs)
        return aabbs, np.zeros(n, np.uint32), np.zeros(n * abi.TR_MAX_LIGHTS_PER_CLUSTER, np.uint32)
    counts, indices = oracle.assign_lights_to_clusters(lights, aabbs, cam.assign_lights())
    return aabbs, counts, indices


def oracle_shade_path(oracle, lut, cam, uniforms, materials, lights, g0, g1, counts, indices, y0=0, y1=None):
    """opaque -> mips -> transmission on the CPU oracle.  Returns dict of every intermediate."""
    pc = cam.push_constants()
    sc = oracle_scene(pc, uniforms, materials, lights, counts, indices)
    out = {}
    if g0 is not None:
        hdr32, hdr16 = oracle.shade_opaque_frame(g0, sc, y0, y1)
    else:
        hdr32 = np.zeros((cam.height, cam.width, 4), np.float32)
        hdr16 = np.zeros((cam.height, cam.width, 4), np.uint16)
    out["opaque32"], out["opaque16"] = hdr32, hdr16
    return out, sc


def gpu_setup(r, lut, uniforms, materials, lights):
    r.set_uniforms(uniforms)
    r.set_materials(materials)
    r.set_lights(lights)
    if lut is not None:
        r.set_ggx_lut(lut)

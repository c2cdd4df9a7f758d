"""ctypes front-end of oracle/_ref/libspvref.so — the reference's SHIPPED SPIR-V modules executed on the CPU
(oracle/spv2c.py + oracle/spv_harness.c).  Same call shapes as oracle/pyoracle.py, so a test can run the C
restatement and the reference's own compiled code side by side.

TEST INFRASTRUCTURE ONLY: importable from tests/ and from scripts under tests/golden/.  `available()` is False on a
machine that has neither /root/reference nor a prebuilt oracle/_ref/libspvref.so.
"""
import ctypes as C

import numpy as np

from oracle import build_ref
from oracle import pyoracle as po
from transmission_renderer_b200 import abi

_LIB = None
_p, _c = po._p, po._c


def available():
    return build_ref.build() is not None


def lib():
    global _LIB
    if _LIB is None:
        po.lib()  # liboracle.so first (libspvref.so links against it)
        so = build_ref.build()
        if so is None:
            raise RuntimeError("neither /root/reference nor a prebuilt oracle/_ref/libspvref.so is present")
        _LIB = C.CDLL(so)
    return _LIB


def frustum_culling(instances, primitives, culling_pc):
    """instance_counts exactly as the module leaves them (one atomic increment per visible instance)."""
    instances = _c(instances, abi.instance)
    primitives = _c(primitives, abi.primitive_info)
    culling_pc = _c(culling_pc, abi.culling_push_constants)
    counts = np.zeros(len(primitives), dtype=np.uint32)
    lib().ref_frustum_culling(_p(instances), C.c_uint32(len(instances)), _p(primitives), C.c_uint32(len(primitives)),
                              _p(culling_pc), _p(counts))
    return counts


def frustum_culling_visible(instances, primitives, culling_pc):
    """Ascending ids of the instances whose invocation incremented its primitive's count."""
    instances = _c(instances, abi.instance)
    primitives = _c(primitives, abi.primitive_info)
    culling_pc = _c(culling_pc, abi.culling_push_constants)
    bits = np.zeros(max(1, len(instances)), dtype=np.uint8)
    lib().ref_frustum_culling_bits(_p(instances), C.c_uint32(len(instances)), _p(primitives), C.c_uint32(len(primitives)),
                                   _p(culling_pc), _p(bits))
    return np.nonzero(bits[: len(instances)])[0].astype(np.uint32)


def demultiplex_draws(primitives, instance_counts):
    primitives = _c(primitives, abi.primitive_info)
    instance_counts = _c(instance_counts, np.uint32)
    n = len(primitives)
    draws = [np.zeros(max(1, n), dtype=abi.draw_indexed_indirect_command) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[d.ctypes.data for d in draws])
    counts = np.zeros(4, dtype=np.uint32)
    lib().ref_demultiplex_draws(_p(primitives), _p(instance_counts), C.c_uint32(n), ptrs, _p(counts))
    return [d[: counts[i]].copy() for i, d in enumerate(draws)], counts


def write_cluster_data(uniforms, wc_pc, nz=16):
    uniforms = _c(uniforms, abi.uniforms)
    wc_pc = _c(wc_pc, abi.write_cluster_data_push_constants)
    nx, ny = int(uniforms["num_clusters"][0, 0]), int(uniforms["num_clusters"][0, 1])
    out = np.zeros(nx * ny * nz, dtype=abi.cluster_aabb)
    lib().ref_write_cluster_data(_p(uniforms), _p(wc_pc), C.c_uint32(nz), _p(out))
    return out


def assign_lights_to_clusters(lights, clusters, al_pc):
    lights = _c(lights, abi.light)
    clusters = _c(clusters, abi.cluster_aabb)
    al_pc = _c(al_pc, abi.assign_lights_push_constants)
    n = len(clusters)
    counts = np.zeros(n, dtype=np.uint32)
    indices = np.zeros(n * abi.TR_MAX_LIGHTS_PER_CLUSTER, dtype=np.uint32)
    lib().ref_assign_lights_to_clusters(_p(lights), C.c_uint32(len(lights)), _p(clusters), C.c_uint32(n), _p(al_pc),
                                        _p(counts), _p(indices))
    return counts, indices


def shade_opaque_frame(gbuffer, scene, y0=0, y1=None):
    pc = scene["push_constants"]
    w, h = int(pc["framebuffer_size"][0, 0]), int(pc["framebuffer_size"][0, 1])
    y1 = h if y1 is None else y1
    g, kg = po._gbuffer_struct(gbuffer, w, h)
    s, ks = po._scene_struct(scene)
    hdr32 = np.zeros((h, w, 4), dtype=np.float32)
    hdr16 = np.zeros((h, w, 4), dtype=np.uint16)
    lib().ref_shade_opaque_frame(C.byref(g), C.byref(s), C.c_uint32(y0), C.c_uint32(y1), _p(hdr32), _p(hdr16), None)
    return hdr32, hdr16


def shade_transmission_frame(gbuffer, scene, levels, lut_rgba8, hdr32, hdr16, y0=0, y1=None):
    pc = scene["push_constants"]
    w, h = int(pc["framebuffer_size"][0, 0]), int(pc["framebuffer_size"][0, 1])
    y1 = h if y1 is None else y1
    g, kg = po._gbuffer_struct(gbuffer, w, h)
    s, ks = po._scene_struct(scene)
    pyr, k1 = po.make_pyramid_struct(levels)
    lut, k2 = po.make_lut_struct(lut_rgba8)
    hdr32 = np.array(hdr32, dtype=np.float32, copy=True)
    hdr16 = np.array(hdr16, dtype=np.uint16, copy=True)
    lib().ref_shade_transmission_frame(C.byref(g), C.byref(s), C.byref(pyr), C.byref(lut), C.c_uint32(y0),
                                       C.c_uint32(y1), _p(hdr32), _p(hdr16))
    return hdr32, hdr16


def tonemap_pixels(rgba, params):
    """fragment_tonemap on (n, 4) linear pixels -> (n, 4) tonemapped linear values (the sRGB encode is the swapchain's)."""
    rgba = _c(rgba, np.float32).reshape(-1, 4)
    params = _c(params, abi.baked_lottes_tonemapper_params)
    out = np.zeros_like(rgba)
    lib().ref_tonemap_pixels(C.c_uint32(len(rgba)), _p(rgba), _p(params), _p(out))
    return out


def vertex_instanced_with_scale(positions, normals, uvs, instance_index, instances, push_constants):
    positions, normals = _c(positions, np.float32).reshape(-1, 3), _c(normals, np.float32).reshape(-1, 3)
    uvs = _c(uvs, np.float32).reshape(-1, 2)
    instance_index = _c(instance_index, np.uint32)
    instances = _c(instances, abi.instance)
    pc = _c(push_constants, abi.push_constants)
    n = len(positions)
    out = dict(clip=np.zeros((n, 4), np.float32), position=np.zeros((n, 3), np.float32), normal=np.zeros((n, 3), np.float32),
               uv=np.zeros((n, 2), np.float32), material_id=np.zeros(n, np.uint32), scale=np.zeros(n, np.float32))
    lib().ref_vertex_instanced_with_scale(C.c_uint32(n), _p(positions), _p(normals), _p(uvs), _p(instance_index),
                                          _p(instances), C.c_uint32(len(instances)), _p(pc), _p(out["clip"]),
                                          _p(out["position"]), _p(out["normal"]), _p(out["uv"]), _p(out["material_id"]),
                                          _p(out["scale"]))
    return out


def vertex_stage(which, positions, normals, uvs, instance_index, instances, push_constants):
    """which = "vertex_instanced" | "depth_pre_pass_instanced" | "depth_pre_pass_vertex_alpha_clip": the outputs that stage has."""
    code = {"vertex_instanced": 1, "depth_pre_pass_instanced": 2, "depth_pre_pass_vertex_alpha_clip": 3}[which]
    positions, normals = _c(positions, np.float32).reshape(-1, 3), _c(normals, np.float32).reshape(-1, 3)
    uvs = _c(uvs, np.float32).reshape(-1, 2)
    instance_index = _c(instance_index, np.uint32)
    instances = _c(instances, abi.instance)
    pc = _c(push_constants, abi.push_constants)
    n = len(positions)
    out = dict(clip=np.zeros((n, 4), np.float32), position=np.zeros((n, 3), np.float32), normal=np.zeros((n, 3), np.float32),
               uv=np.zeros((n, 2), np.float32), material_id=np.zeros(n, np.uint32))
    rc = lib().ref_vertex_stage(C.c_uint32(code), C.c_uint32(n), _p(positions), _p(normals), _p(uvs), _p(instance_index),
                                _p(instances), C.c_uint32(len(instances)), _p(pc), _p(out["clip"]), _p(out["position"]),
                                _p(out["normal"]), _p(out["uv"]), _p(out["material_id"]))
    assert rc == 0
    keep = {1: ("clip", "position", "normal", "uv", "material_id"), 2: ("clip",), 3: ("clip", "uv", "material_id")}[code]
    return {k: out[k] for k in keep}


def alpha_clip(uvs, duv, material_id, scene):
    uvs = _c(uvs, np.float32).reshape(-1, 2)
    duv = _c(duv, np.float32).reshape(-1, 4) if duv is not None else None
    material_id = _c(material_id, np.uint32)
    s, ks = po._scene_struct(scene)
    out = np.zeros(len(uvs), np.uint8)
    lib().ref_alpha_clip(C.c_uint32(len(uvs)), _p(uvs), _p(duv), _p(material_id), C.byref(s), _p(out))
    return out


class ShadePathRunner(po.ShadePathRunner):
    """The oracle's shade-path runner with both fragment stages executed by the reference's own compiled modules
    (fragment.spv / fragment_transmission.spv, or their ray-tracing builds when the scene carries an acceleration structure):
    same buffers, same frame drivers (G-buffer decode, OpenMP over rows), same call shapes.  The mip chain (vkCmdBlitImage in the
    reference) and the sRGB8 encode stay the oracle's.  bench.py's `--impl reference` / `cpu_baseline` legs time this when
    oracle/_ref/libspvref.so is present: the reference's shader code itself on the host cores."""

    def opaque(self, y0, y1):
        lib().ref_shade_opaque_frame(C.byref(self._g0), C.byref(self._s), C.c_uint32(y0), C.c_uint32(y1), _p(self.hdr32),
                                     _p(self.hdr16), _p(self.levels[0]))

    def transmission(self, y0, y1):
        if self._g1 is not None:
            lib().ref_shade_transmission_frame(C.byref(self._g1), C.byref(self._s), C.byref(self._pyr), C.byref(self._lut),
                                               C.c_uint32(y0), C.c_uint32(y1), _p(self.hdr32), _p(self.hdr16))

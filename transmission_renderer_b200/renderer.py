"""`Renderer`: Python mirror of the C ABI (include/tr_abi.h), one method per `tr_*` entry point.

It plays the part of the reference's frame recorder `record()` (src/main.rs:1551-2263) for
tests and benchmarks: same pass names, same argument structs, same error behaviour (a
non-zero status raises `TrError` carrying `tr_last_error()`, like `anyhow::Result` + `?`).
All arithmetic happens in libtr.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import abi


def _lib():
    from . import lib
    return lib()


def _check(status):
    from . import _check as chk
    chk(status)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class _Config(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("device", C.c_int32), ("band_y0", C.c_uint32),
                ("band_y1", C.c_uint32), ("flags", C.c_uint32)]


class _Planes(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("normal", C.c_void_p), ("uv", C.c_void_p), ("material_id", C.c_void_p),
                ("scale", C.c_void_p), ("position", C.c_void_p), ("duv", C.c_void_p), ("ddepth", C.c_void_p)]


_PLANES = ("depth", "normal", "uv", "material_id", "scale", "position", "duv", "ddepth")


class Renderer:
    def __init__(self, width, height, device=0, band=None, f32_debug=False):
        self._ctx = C.c_void_p()
        self.width, self.height = int(width), int(height)
        y0, y1 = band if band is not None else (0, 0)
        cfg = _Config(self.width, self.height, device, y0, y1, abi.TR_FLAG_HDR_F32_DEBUG if f32_debug else 0)
        _check(_lib().tr_create(C.byref(cfg), C.byref(self._ctx)))
        self.f32_debug = f32_debug
        self._keep = []

    # -- lifecycle -------------------------------------------------------------------
    def close(self):
        if self._ctx:
            _lib().tr_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def resize(self, width, height):
        _check(_lib().tr_resize(self._ctx, C.c_uint32(width), C.c_uint32(height)))
        self.width, self.height = int(width), int(height)

    def set_band(self, y0, y1):
        _check(_lib().tr_set_band(self._ctx, C.c_uint32(y0), C.c_uint32(y1)))

    def set_stream(self, cuda_stream_handle):
        _check(_lib().tr_set_stream(self._ctx, C.c_void_p(cuda_stream_handle)))

    def sync(self):
        _check(_lib().tr_sync(self._ctx))

    # -- uploads ------------------------------------------------------------------------
    def set_instances(self, instances):
        a = _c(instances, abi.instance)
        _check(_lib().tr_set_instances(self._ctx, _p(a), C.c_uint32(len(a))))

    def set_primitives(self, primitives):
        a = _c(primitives, abi.primitive_info)
        _check(_lib().tr_set_primitives(self._ctx, _p(a), C.c_uint32(len(a))))

    def set_materials(self, materials):
        a = _c(materials, abi.material_info)
        _check(_lib().tr_set_materials(self._ctx, _p(a), C.c_uint32(len(a))))

    def set_lights(self, lights):
        a = _c(lights, abi.light)
        _check(_lib().tr_set_lights(self._ctx, _p(a) if len(a) else None, C.c_uint32(len(a))))

    def set_uniforms(self, uniforms):
        a = _c(uniforms, abi.uniforms)
        _check(_lib().tr_set_uniforms(self._ctx, _p(a)))

    def set_ggx_lut(self, rgba8):
        a = _c(rgba8, np.uint8)
        _check(_lib().tr_set_ggx_lut(self._ctx, _p(a), C.c_uint32(a.shape[1]), C.c_uint32(a.shape[0])))

    def set_texture(self, index, levels, srgb):
        """levels: list of (h, w, 4) uint8 mips, level 0 first (scenes.make_mips)."""
        lv = [_c(l, np.uint8) for l in levels]
        ptrs = (C.c_void_p * len(lv))(*[l.ctypes.data for l in lv])
        _check(_lib().tr_set_texture(self._ctx, C.c_uint32(index), ptrs, C.c_uint32(len(lv)), C.c_uint32(lv[0].shape[1]),
                                     C.c_uint32(lv[0].shape[0]), C.c_int32(1 if srgb else 0)))

    def set_textures(self, textures):
        for i, t in enumerate(textures):
            self.set_texture(i, t["levels"], t["srgb"])

    def set_mesh(self, positions, normals, uvs, indices):
        p, n, u, i = _c(positions, np.float32), _c(normals, np.float32), _c(uvs, np.float32), _c(indices, np.uint32)
        _check(_lib().tr_set_mesh(self._ctx, _p(p), _p(n), _p(u), C.c_uint32(len(p)), _p(i), C.c_uint32(i.size)))

    # -- ray-queried shadows (`--ray-tracing`, src/main.rs:577-658) ---------------------------
    def build_acceleration_structures(self):
        """Bottom-level structures of every primitive + the top level over the instances with draw_buffer_index < 2.
        Returns the handle to put into PushConstants.acceleration_structure_address (src/main.rs:856-859)."""
        a = C.c_uint64(0)
        _check(_lib().tr_build_acceleration_structures(self._ctx, C.byref(a)))
        return a.value

    def update_top_level_acceleration_structure(self):
        a = C.c_uint64(0)
        _check(_lib().tr_update_top_level_acceleration_structure(self._ctx, C.byref(a)))
        return a.value

    def trace_shadow_rays(self, origins, directions, t_max):
        """(n,3), (n,3), (n,) -> (n,) uint8: 1 = lit, 0 = occluded (trace_shadow_ray, lighting.rs:97-125)."""
        o, d = _c(origins, np.float32).reshape(-1, 3), _c(directions, np.float32).reshape(-1, 3)
        t = _c(t_max, np.float32).reshape(-1)
        out = np.zeros(len(o), np.uint8)
        _check(_lib().tr_trace_shadow_rays(self._ctx, C.c_uint32(len(o)), _p(o), _p(d), _p(t), _p(out)))
        return out

    def read_shadow_mask(self, layer):
        """(5, h, w) uint32 occluded-ray bits of the last shading pass of `layer` (band rows)."""
        out = np.zeros((5, self.height, self.width), np.uint32)
        _check(_lib().tr_read_shadow_mask(self._ctx, C.c_int32(layer), _p(out)))
        return out

    # -- passes, in record() order --------------------------------------------------------
    def cull(self, culling_pc):
        _check(_lib().tr_cull(self._ctx, _p(_c(culling_pc, abi.culling_push_constants))))

    def build_clusters(self, wc_pc):
        _check(_lib().tr_build_clusters(self._ctx, _p(_c(wc_pc, abi.write_cluster_data_push_constants))))

    def assign_lights(self, al_pc):
        _check(_lib().tr_assign_lights(self._ctx, _p(_c(al_pc, abi.assign_lights_push_constants))))

    def visibility(self, pc):
        _check(_lib().tr_visibility(self._ctx, _p(_c(pc, abi.push_constants))))

    def shade_opaque(self, pc):
        _check(_lib().tr_shade_opaque(self._ctx, _p(_c(pc, abi.push_constants))))

    def allgather_opaque(self):
        _check(_lib().tr_allgather_opaque(self._ctx))

    def generate_mips(self):
        _check(_lib().tr_generate_mips(self._ctx))

    def shade_transmission(self, pc):
        _check(_lib().tr_shade_transmission(self._ctx, _p(_c(pc, abi.push_constants))))

    def tonemap(self, params):
        _check(_lib().tr_tonemap(self._ctx, _p(_c(params, abi.baked_lottes_tonemapper_params))))

    def frame(self, frame_params):
        _check(_lib().tr_frame(self._ctx, _p(_c(frame_params, abi.frame_params))))

    def begin_frame(self):
        """Start of a frame for the per-pass timers when the passes are called one by one (tr_frame does it itself)."""
        _check(_lib().tr_begin_frame(self._ctx))

    def enable_timing(self, on=True):
        _check(_lib().tr_enable_timing(self._ctx, C.c_int32(1 if on else 0)))

    def frame_times(self):
        t = np.zeros(1, dtype=abi.frame_times)
        _check(_lib().tr_read_frame_times(self._ctx, _p(t)))
        return {k: float(t[k][0]) for k in t.dtype.names}

    def pass_totals(self):
        """(sum of per-pass ms over the frames since enable_timing / the last call, number of frames)."""
        t = np.zeros(1, dtype=abi.frame_times)
        n = C.c_uint32(0)
        _check(_lib().tr_read_pass_totals(self._ctx, _p(t), C.byref(n)))
        return {k: float(t[k][0]) for k in t.dtype.names}, n.value

    # -- parity hooks ------------------------------------------------------------------------
    def set_gbuffer(self, layer, gbuffer):
        npx = self.width * self.height
        keep = dict(depth=_c(gbuffer["depth"], np.float32), normal=_c(gbuffer["normal"], np.float32),
                    material_id=_c(gbuffer["material_id"], np.uint32))
        assert keep["depth"].size == npx and keep["normal"].size == npx * 3 and keep["material_id"].size == npx
        for k, dt, mult in (("uv", np.float32, 2), ("scale", np.float32, 1), ("position", np.float32, 3),
                            ("duv", np.float32, 4), ("ddepth", np.float32, 2)):
            v = gbuffer.get(k)
            keep[k] = None if v is None else _c(v, dt)
            assert keep[k] is None or keep[k].size == npx * mult
        planes = _Planes(*[None if keep[k] is None else keep[k].ctypes.data for k in _PLANES])
        _check(_lib().tr_set_gbuffer(self._ctx, C.c_int32(layer), C.byref(planes)))
        self.sync()

    def read_gbuffer(self, layer, with_position=False, derivatives=False):
        h, w = self.height, self.width
        g = dict(depth=np.zeros((h, w), np.float32), normal=np.zeros((h, w, 3), np.float32),
                 uv=np.zeros((h, w, 2), np.float32), material_id=np.zeros((h, w), np.uint32),
                 scale=np.zeros((h, w), np.float32), position=np.zeros((h, w, 3), np.float32) if with_position else None,
                 duv=np.zeros((h, w, 4), np.float32) if derivatives else None,
                 ddepth=np.zeros((h, w, 2), np.float32) if derivatives else None)
        planes = _Planes(*[None if g[k] is None else g[k].ctypes.data for k in _PLANES])
        _check(_lib().tr_read_gbuffer(self._ctx, C.c_int32(layer), C.byref(planes)))
        return g

    def set_opaque_frame(self, rgba16f_bits):
        a = _c(rgba16f_bits, np.uint16)
        assert a.size == self.width * self.height * 4
        _check(_lib().tr_set_opaque_frame(self._ctx, _p(a)))
        self.sync()

    def set_hdr(self, rgba16f_bits):
        a = _c(rgba16f_bits, np.uint16)
        assert a.size == self.width * self.height * 4
        _check(_lib().tr_set_hdr(self._ctx, _p(a)))
        self.sync()

    def set_cluster_lights(self, counts, indices):
        a, b = _c(counts, np.uint32), _c(indices, np.uint32)
        _check(_lib().tr_set_cluster_lights(self._ctx, _p(a), _p(b)))
        self.sync()

    def read_visible_instances(self):
        n = C.c_uint32(0)
        _check(_lib().tr_read_visible_instances(self._ctx, None, C.c_uint32(0), C.byref(n)))
        ids = np.zeros(max(1, n.value), dtype=np.uint32)
        _check(_lib().tr_read_visible_instances(self._ctx, _p(ids), C.c_uint32(len(ids)), C.byref(n)))
        return ids[: n.value].copy()

    def read_instance_counts(self, n_primitives):
        a = np.zeros(n_primitives, dtype=np.uint32)
        _check(_lib().tr_read_instance_counts(self._ctx, _p(a), C.c_uint32(n_primitives)))
        return a

    def read_draws(self, bucket):
        n = C.c_uint32(0)
        _check(_lib().tr_read_draws(self._ctx, C.c_uint32(bucket), None, C.c_uint32(0), C.byref(n)))
        cmds = np.zeros(max(1, n.value), dtype=abi.draw_indexed_indirect_command)
        _check(_lib().tr_read_draws(self._ctx, C.c_uint32(bucket), _p(cmds), C.c_uint32(len(cmds)), C.byref(n)))
        return cmds[: n.value].copy()

    def read_cluster_aabbs(self, n_clusters):
        a = np.zeros(n_clusters, dtype=abi.cluster_aabb)
        _check(_lib().tr_read_cluster_aabbs(self._ctx, _p(a), C.c_uint32(n_clusters)))
        return a

    def read_cluster_lights(self, n_clusters):
        counts = np.zeros(n_clusters, dtype=np.uint32)
        indices = np.zeros(n_clusters * abi.TR_MAX_LIGHTS_PER_CLUSTER, dtype=np.uint32)
        _check(_lib().tr_read_cluster_lights(self._ctx, _p(counts), _p(indices)))
        return counts, indices

    def read_hdr(self):
        a = np.zeros((self.height, self.width, 4), dtype=np.uint16)
        _check(_lib().tr_read_hdr(self._ctx, _p(a)))
        return a

    def read_hdr_f32(self):
        a = np.zeros((self.height, self.width, 4), dtype=np.float32)
        _check(_lib().tr_read_hdr_f32(self._ctx, _p(a)))
        return a

    def mip_levels(self):
        n = C.c_uint32(0)
        _check(_lib().tr_mip_levels(self._ctx, C.byref(n)))
        return n.value

    def read_pyramid_level(self, level):
        w, h = C.c_uint32(0), C.c_uint32(0)
        _check(_lib().tr_read_pyramid_level(self._ctx, C.c_uint32(level), None, C.byref(w), C.byref(h)))
        a = np.zeros((h.value, w.value, 4), dtype=np.uint16)
        _check(_lib().tr_read_pyramid_level(self._ctx, C.c_uint32(level), _p(a), C.byref(w), C.byref(h)))
        return a

    def read_srgb8(self, out=None):
        a = np.zeros((self.height, self.width, 4), dtype=np.uint8) if out is None else out
        _check(_lib().tr_read_srgb8(self._ctx, _p(a)))
        return a

    # -- glam-pbr contracts --------------------------------------------------------------------
    def read_srgb8_async(self, out):
        """Enqueue the band read-back behind the frame just recorded; `out` (H, W, 4) uint8, ideally pinned."""
        assert out.dtype == np.uint8 and out.size == self.width * self.height * 4 and out.flags.c_contiguous
        _check(_lib().tr_read_srgb8_async(self._ctx, _p(out)))

    def wait_readback(self):
        _check(_lib().tr_wait_readback(self._ctx))

    def eval_basic_brdf(self, params):
        a = _c(params, abi.basic_brdf_params)
        out = np.zeros(len(a), dtype=abi.brdf_result)
        _check(_lib().tr_eval_basic_brdf(self._ctx, C.c_uint32(len(a)), _p(a), _p(out)))
        return out

    def eval_transmission_btdf(self, params):
        a = _c(params, abi.transmission_btdf_params)
        out = np.zeros((len(a), 3), dtype=np.float32)
        _check(_lib().tr_eval_transmission_btdf(self._ctx, C.c_uint32(len(a)), _p(a), _p(out)))
        return out

    def eval_point_light(self, params):
        """The frame kernels' light loop for one (pixel, point light) pair per element."""
        a = _c(params, abi.point_light_params)
        out = np.zeros(len(a), dtype=abi.point_light_result)
        _check(_lib().tr_eval_point_light(self._ctx, C.c_uint32(len(a)), _p(a), _p(out)))
        return out

    def eval_ibl_volume_refraction(self, proj_view, params):
        a = _c(params, abi.ibl_volume_refraction_params)
        pv = _c(np.asarray(proj_view, dtype=np.float32).T, np.float32)
        out = np.zeros((len(a), 3), dtype=np.float32)
        _check(_lib().tr_eval_ibl_volume_refraction(self._ctx, C.c_uint32(len(a)), _p(pv), _p(a), _p(out)))
        return out

    # -- multi-GPU ---------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        _check(_lib().tr_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, n_ranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(_lib().tr_comm_init(self._ctx, buf, C.c_int32(rank), C.c_int32(n_ranks)))

    def set_bands(self, bounds):
        """n_ranks + 1 ascending row boundaries, the same on every rank (tr_set_bands)."""
        a = _c(bounds, np.uint32)
        _check(_lib().tr_set_bands(self._ctx, _p(a), C.c_uint32(len(a))))

    def comm_destroy(self):
        _check(_lib().tr_comm_destroy(self._ctx))

    def peer_export(self):
        buf = (C.c_uint8 * 64)()
        _check(_lib().tr_peer_export(self._ctx, buf))
        return bytes(buf)

    def peer_attach(self, rank, n_ranks, handles):
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(_lib().tr_peer_attach(self._ctx, C.c_int32(rank), C.c_int32(n_ranks), buf))

    def device_buffer(self, what):
        ptr, nbytes = C.c_void_p(), C.c_size_t(0)
        _check(_lib().tr_device_buffer(self._ctx, C.c_int32(what), C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    # -- diagnostics -----------------------------------------------------------------------------------
    @staticmethod
    def launch_count():
        n = C.c_uint64(0)
        _check(_lib().tr_launch_count(C.byref(n)))
        return n.value

    def raster_stats(self, reset=True):
        out = (C.c_uint64 * 4)()
        _check(_lib().tr_raster_stats(self._ctx, out, C.c_int32(1 if reset else 0)))
        return dict(box_pixels=out[0], box_pixels_after_hiz=out[1], exact_evaluations=out[2], span_pixels=out[3])

    def measure_fp32_peak(self):
        v = C.c_float(0)
        _check(_lib().tr_measure_fp32_peak(self._ctx, C.byref(v)))
        return v.value

    def measure_hbm_copy(self):
        v = C.c_float(0)
        _check(_lib().tr_measure_hbm_copy(self._ctx, C.byref(v)))
        return v.value

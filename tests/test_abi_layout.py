"""Struct layouts: numpy mirrors vs SURVEY.md Appendix A (shared-structs/src/lib.rs)."""
import numpy as np

from transmission_renderer_b200 import abi

EXPECT = {
    "push_constants": (96, {"proj_view": 0, "view_position": 64, "framebuffer_size": 80,
                            "acceleration_structure_address": 88}),
    "uniforms": (96, {"z_near": 0, "z_far": 4, "scale": 8, "bias": 12, "num_depth_slices": 16, "sun_dir": 32,
                      "sun_intensity": 48, "cluster_size_in_pixels": 64, "num_clusters": 72, "debug_clusters": 80,
                      "ggx_lut_texture_index": 84}),
    "light": (48, {"position_and_spotlight_epsilon": 0, "colour_emission_and_falloff_distance_sq": 16,
                   "spotlight_direction_and_outer_angle": 32}),
    "material_info": (160, {"textures": 0, "metallic_factor": 36, "roughness_factor": 40, "alpha_clipping_cutoff": 44,
                            "diffuse_factor": 48, "emissive_factor": 64, "normal_map_scale": 80,
                            "occlusion_strength": 84, "index_of_refraction": 88, "transmission_factor": 92,
                            "thickness_factor": 96, "attenuation_distance": 100, "attenuation_colour": 112,
                            "specular_factor": 128, "specular_colour_factor": 144}),
    "instance": (48, {"translation_and_scale": 0, "rotation": 16, "primitive_id": 32, "material_id": 36}),
    "primitive_info": (32, {"packed_bounding_sphere": 0, "draw_buffer_index": 16, "index_count": 20,
                            "first_index": 24, "first_instance": 28}),
    "culling_push_constants": (96, {"view": 0, "frustum_x_xz": 64, "frustum_y_yz": 72, "z_near": 80}),
    "cluster_aabb": (32, {"min": 0, "max": 16}),
    "write_cluster_data_push_constants": (80, {"inverse_perspective": 0, "screen_dimensions": 64}),
    "assign_lights_push_constants": (80, {"view_matrix": 0, "view_rotation": 64}),
    "draw_indexed_indirect_command": (20, {"index_count": 0, "instance_count": 4, "first_index": 8,
                                           "vertex_offset": 12, "first_instance": 16}),
    "baked_lottes_tonemapper_params": (28, {"a": 0, "cross_saturation": 24}),
}


def test_layouts():
    for name, (size, offsets) in EXPECT.items():
        dt = getattr(abi, name)
        assert dt.itemsize == size, name
        for field, off in offsets.items():
            assert dt.fields[field][1] == off, (name, field)


def test_header_compiles_as_c_and_cpp(tmp_path):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "tr_abi.h"\nint main(void){return 0;}\n')
    inc = os.path.join(root, "include")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", inc, "-c", str(src), "-o", str(tmp_path / "a.o")])
    subprocess.check_call(["g++", "-std=c++17", "-x", "c++", "-Wall", "-Werror", "-I", inc, "-c", str(src), "-o",
                           str(tmp_path / "b.o")])


def test_default_material_matches_loader_defaults():
    m = abi.default_material()
    assert m["index_of_refraction"][0] == np.float32(1.5)      # model_loading.rs:306
    assert np.isinf(m["attenuation_distance"][0])               # model_loading.rs:318
    assert m["alpha_clipping_cutoff"][0] == np.float32(0.5)     # model_loading.rs:295
    assert (m["textures"] == -1).all()


# ---- against the layouts the reference's own compiled shaders declare -----------------------------------------------
# tests/golden/spv_layouts.json = OpMemberDecorate Offset / OpDecorate ArrayStride of every buffer and push-constant
# block in /root/reference/compiled-shaders/normal/*.spv (tests/golden/make_spv_layouts.py).  (module, set, binding) ->
# the struct of include/tr_abi.h bound there (shader/src/lib.rs entry-point signatures); None = push constants.
SPV_BINDINGS = {
    ("frustum_culling", 0, 0): "primitive_info", ("frustum_culling", 1, 0): "instance",
    ("frustum_culling", None, None): "culling_push_constants",
    ("demultiplex_draws", 0, 0): "primitive_info", ("demultiplex_draws", 0, 3): "draw_indexed_indirect_command",
    ("demultiplex_draws", 0, 4): "draw_indexed_indirect_command", ("demultiplex_draws", 0, 5): "draw_indexed_indirect_command",
    ("demultiplex_draws", 0, 6): "draw_indexed_indirect_command",
    ("write_cluster_data", 0, 3): "uniforms", ("write_cluster_data", 1, 0): "cluster_aabb",
    ("write_cluster_data", None, None): "write_cluster_data_push_constants",
    ("assign_lights_to_clusters", 0, 0): "light", ("assign_lights_to_clusters", 1, 0): "cluster_aabb",
    ("assign_lights_to_clusters", None, None): "assign_lights_push_constants",
    ("fragment", 0, 2): "material_info", ("fragment", 0, 3): "uniforms", ("fragment", 2, 0): "light",
    ("fragment", None, None): "push_constants",
    ("fragment_transmission", 0, 2): "material_info", ("fragment_transmission", 0, 3): "uniforms",
    ("fragment_transmission", 2, 0): "light", ("fragment_transmission", None, None): "push_constants",
    ("fragment_tonemap", None, None): "baked_lottes_tonemapper_params",
    ("vertex_instanced", 1, 0): "instance", ("vertex_instanced", None, None): "push_constants",
    ("vertex_instanced_with_scale", 1, 0): "instance", ("vertex_instanced_with_scale", None, None): "push_constants",
    ("depth_pre_pass_instanced", 1, 0): "instance", ("depth_pre_pass_alpha_clip", 0, 2): "material_info",
}


def _spv_leaves(t, base=0):
    """[(offset, kind)] of every 32-bit scalar a SPIR-V type description lays out."""
    if isinstance(t, str):
        kind, _, n = t.partition("x")
        return [(base + 4 * k, kind[0]) for k in range(int(n) if n else 1)]
    if "struct" in t:
        out = []
        for m in t["struct"]:
            out += _spv_leaves(m["type"], base + m["offset"])
        return out
    if "array" in t:
        out = []
        for k in range(t["len"]):
            out += _spv_leaves(t["array"], base + k * t["stride"])
        return out
    raise AssertionError(t)


def _np_leaves(dt, base=0):
    """{offset of each 4-byte word: kind} of a numpy struct dtype (u8 covers two words)."""
    out = {}
    if dt.fields is None:
        sub, shape = (dt.subdtype if dt.subdtype else (dt, ()))
        count = int(np.prod(shape)) if shape else 1
        for k in range(count):
            for wd in range(sub.itemsize // 4):
                out[base + k * sub.itemsize + wd * 4] = sub.kind
        return out
    for name in dt.names:
        fdt, off = dt.fields[name][:2]
        out.update(_np_leaves(fdt, base + off))
    return out


def test_layouts_match_the_shipped_spirv():
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spv = json.load(open(os.path.join(root, "tests", "golden", "spv_layouts.json")))
    checked = 0
    for (module, s, b), name in SPV_BINDINGS.items():
        dt = getattr(abi, name)
        block = next(i for i in spv[module]["interface"] if i["set"] == s and i["binding"] == b)["type"]
        inner = block["struct"][0]["type"]           # rust-gpu wraps every binding in a one-member Block struct
        if isinstance(inner, dict) and "rtarray" in inner:
            assert inner["stride"] == dt.itemsize, (module, name, "ArrayStride")
            inner = inner["rtarray"]
        have = _np_leaves(dt)
        kinds = {"f": "f", "u": "u", "i": "i"}
        for off, kind in _spv_leaves(inner):
            assert off in have, (module, name, off)
            assert have[off] == kinds[kind], (module, name, off, kind, have[off])
            checked += 1
        assert max(o for o, _ in _spv_leaves(inner)) + 4 <= dt.itemsize
    assert checked > 300

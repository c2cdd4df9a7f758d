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

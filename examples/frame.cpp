// examples/frame.cpp — one frame through the C++ host mirror (include/tr_host.hpp): what the reference's main() does between
// loading a model and presenting a frame (src/main.rs:377-552 set-up, :1551-2263 record), with a procedural mesh.
//
//   g++ -std=c++17 -I include examples/frame.cpp -L transmission_renderer_b200 -ltr -Wl,-rpath,$PWD/transmission_renderer_b200 -o frame_cpp
//
// Without a CUDA device tr_create fails, tr::Error carries the library's message, and the program exits with status 3:
// there is no CPU path.
#include <cstdio>

#include "tr_host.hpp"

int main() {
    const uint32_t w = 640, h = 360;
    try {
        tr::Renderer r(w, h);
        // a ground quad and a pyramid above it; every triangle with both windings so that the example needs no care
        std::vector<float> pos = {-4, 0, -8, 4, 0, -8, 4, 0, 0, -4, 0, 0, 0, 2.5f, -4, -1, 0.5f, -3, 1, 0.5f, -3, 0, 0.5f, -5};
        std::vector<float> nrm = {0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, -1, 0, 1, 1, 0, 1, 0, 0, -1};
        std::vector<float> uv = {0, 0, 1, 0, 1, 1, 0, 1, 0.5f, 0.5f, 0, 0, 1, 0, 0.5f, 1};
        std::vector<uint32_t> one_side = {0, 1, 2, 0, 2, 3, 4, 5, 6, 4, 6, 7, 4, 7, 5};
        std::vector<uint32_t> idx;
        for (size_t i = 0; i < one_side.size(); i += 3)
            for (uint32_t k : {one_side[i], one_side[i + 1], one_side[i + 2], one_side[i], one_side[i + 2], one_side[i + 1]}) idx.push_back(k);
        r.set_mesh(pos, nrm, uv, idx);

        tr_primitive_info prim = {};
        prim.packed_bounding_sphere.z = -4.0f;
        prim.packed_bounding_sphere.w = 8.0f;
        prim.index_count = (uint32_t)idx.size();
        r.set_primitives({prim});
        tr_instance inst = {};
        inst.transform.translation_and_scale.w = 1.0f;
        inst.transform.rotation.w = 1.0f;
        r.set_instances({inst});
        tr_material_info mat = {};
        mat.textures = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
        mat.roughness_factor = 0.5f;
        mat.alpha_clipping_cutoff = 0.5f;
        mat.diffuse_factor.x = 0.8f; mat.diffuse_factor.y = 0.3f; mat.diffuse_factor.z = 0.2f; mat.diffuse_factor.w = 1.0f;
        mat.index_of_refraction = 1.5f;
        mat.attenuation_distance = INFINITY;
        mat.specular_factor = 1.0f;
        mat.specular_colour_factor.x = mat.specular_colour_factor.y = mat.specular_colour_factor.z = 1.0f;
        r.set_materials({mat});
        // the reference's two default lights, src/main.rs:450-453
        r.set_lights({tr::light_new_point({0.0f, 0.8f, 0.0f}, {1.0f, 0.0f, 0.0f}, 5.0f), tr::light_new_point({8.0f, 0.8f, 0.0f}, {0.0f, 1.0f, 0.0f}, 10.0f)});
        r.set_uniforms(tr::make_uniforms(w, h));
        r.set_ggx_lut(std::vector<uint8_t>(4 * 4 * 4, 128), 4, 4);   // the reference loads ggx_lut.png (src/main.rs:295-330); a flat stand-in
        r.build_clusters();

        const tr::Camera cam = tr::yaw_pitch_camera({0.0f, 3.0f, 1.0f}, 0.0, -15.0);   // src/main.rs:514-516
        r.frame(tr::make_frame_params(cam, w, h, tr::default_tonemap_params()));
        const std::vector<uint8_t> rgba = r.read_srgb8();
        size_t lit = 0;
        for (size_t i = 0; i < (size_t)w * h; i++) lit += (rgba[i * 4] | rgba[i * 4 + 1] | rgba[i * 4 + 2]) != 0;
        std::printf("%s: frame %ux%u, %zu lit pixels, %zu visible instance(s)\n", tr_version(), w, h, lit, r.read_visible_instances(1).size());
    } catch (const tr::Error& e) {
        std::fprintf(stderr, "%s (status %d)\n", e.what(), e.status());
        return 3;
    }
    return 0;
}

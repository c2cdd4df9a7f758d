// tests/host_mirror_dump.cpp — writes the structs include/tr_host.hpp builds for one camera to stdout (binary, fixed order);
// tests/test_host_mirror.py compares them with transmission_renderer_b200/host.py.  No GPU, no libtr.so call.
//   host_mirror_dump width height px py pz yaw_deg pitch_deg
#include <cstdio>
#include <cstdlib>

#include "tr_host.hpp"

template <typename T>
static void put(const T& v) { std::fwrite(&v, sizeof v, 1, stdout); }

int main(int argc, char** argv) {
    if (argc != 8) return 2;
    const uint32_t w = (uint32_t)std::atoi(argv[1]), h = (uint32_t)std::atoi(argv[2]);
    const tr::Vec3 pos = {(float)std::atof(argv[3]), (float)std::atof(argv[4]), (float)std::atof(argv[5])};
    const tr::Camera cam = tr::yaw_pitch_camera(pos, std::atof(argv[6]), std::atof(argv[7]));
    put(tr::make_frame_params(cam, w, h, tr::default_tonemap_params(), TR_FRAME_SKIP_TONEMAP, 0x1234));   // 304 bytes
    put(tr::make_uniforms(w, h));                                                                         // 96
    put(tr::make_write_cluster_data_push_constants(tr::perspective_matrix_reversed(w, h), w, h));         // 80
    put(tr::light_new_point({0.5f, 3.0f, 1.5f}, {1.0f, 0.8f, 0.6f}, 8.0f));                               // 48
    put(tr::light_new_spot({-8.0f, 9.0f, -14.0f}, {1.0f, 0.9f, 0.8f}, 40.0f, {0.0f, -1.0f, 0.0f}, 0.3f, 0.5f));   // 48
    const uint32_t misc[8] = {tr::mip_levels_for_size(w, h), tr::dispatch_count(10000, 64), tr::band_rows(h, 3, 8).first, tr::band_rows(h, 3, 8).second,
                              tr::NUM_CLUSTERS, tr::dispatch_count(0, 64), tr::mip_levels_for_size(7680, 4320), tr::band_rows(h, 7, 8).second};
    put(misc);
    return 0;
}

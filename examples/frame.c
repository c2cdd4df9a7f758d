/* examples/frame.c — the drop-in boundary from plain C: what a host that links libtr.so does for one frame.
 *
 *   gcc -std=c11 -I include examples/frame.c -L transmission_renderer_b200 -ltr -Wl,-rpath,$PWD/transmission_renderer_b200 -lm -o frame
 *
 * The calls are the ones INTEGRATION.md maps onto the reference's `record()` (src/main.rs:1551-2263).  There is no CPU
 * fallback: without a CUDA device tr_create fails and the program says so and exits with status 3 (that, and that the
 * header is valid C11, is what tests/test_abi_symbols.py checks; the parity tests drive the same calls through ctypes).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tr_abi.h"

#define CHECK(call)                                                               \
    do {                                                                          \
        int32_t st_ = (call);                                                     \
        if (st_ != TR_OK) {                                                       \
            fprintf(stderr, "%s -> %d: %s\n", #call, st_, tr_last_error());       \
            return 3;                                                             \
        }                                                                         \
    } while (0)

static void identity(tr_mat4* m) {
    memset(m, 0, sizeof(*m));
    m->col[0].x = m->col[1].y = m->col[2].z = m->col[3].w = 1.0f;
}

int main(void) {
    const uint32_t w = 256, h = 144;
    tr_config cfg = {w, h, 0, 0, 0, 0};
    tr_ctx* ctx = NULL;
    printf("%s\n", tr_version());
    CHECK(tr_create(&cfg, &ctx));

    /* one triangle in clip space: proj_view = identity, reversed-Z depth 0.5 */
    const float pos[9] = {-0.8f, -0.8f, 0.5f, 0.8f, -0.8f, 0.5f, 0.0f, 0.8f, 0.5f};
    const float nrm[9] = {0, 0, 1, 0, 0, 1, 0, 0, 1};
    const float uv[6] = {0, 0, 1, 0, 0.5f, 1};
    const uint32_t idx[6] = {0, 1, 2, 0, 2, 1};   /* both windings: whichever faces the camera survives back-face culling */
    tr_primitive_info prim;
    memset(&prim, 0, sizeof(prim));
    prim.packed_bounding_sphere.w = 2.0f;
    prim.index_count = 6;
    tr_instance inst;
    memset(&inst, 0, sizeof(inst));
    inst.transform.translation_and_scale.w = 1.0f;
    inst.transform.rotation.w = 1.0f;
    tr_material_info mat;
    memset(&mat, 0, sizeof(mat));
    mat.textures.diffuse = mat.textures.metallic_roughness = mat.textures.normal_map = mat.textures.emissive = -1;
    mat.textures.occlusion = mat.textures.transmission = mat.textures.thickness = mat.textures.specular = mat.textures.specular_colour = -1;
    mat.roughness_factor = 0.5f;
    mat.alpha_clipping_cutoff = 0.5f;
    mat.diffuse_factor.x = 0.8f; mat.diffuse_factor.y = 0.3f; mat.diffuse_factor.z = 0.2f; mat.diffuse_factor.w = 1.0f;
    mat.index_of_refraction = 1.5f;
    mat.attenuation_distance = INFINITY;
    mat.specular_factor = 1.0f;
    mat.specular_colour_factor.x = mat.specular_colour_factor.y = mat.specular_colour_factor.z = 1.0f;
    tr_uniforms u;
    memset(&u, 0, sizeof(u));
    u.light_clustering_coefficients.z_near = 0.01f;
    u.light_clustering_coefficients.z_far = 500.0f;
    u.light_clustering_coefficients.num_depth_slices = 24;
    u.light_clustering_coefficients.scale = 24.0f / log2f(500.0f / 0.01f);
    u.light_clustering_coefficients.bias = -(24.0f * log2f(0.01f) / log2f(500.0f / 0.01f));
    u.sun_dir.z = 1.0f;
    u.sun_intensity.x = u.sun_intensity.y = u.sun_intensity.z = 3.0f;
    u.num_clusters.x = 12; u.num_clusters.y = 8;
    u.cluster_size_in_pixels.x = (float)w / 12.0f; u.cluster_size_in_pixels.y = (float)h / 8.0f;

    CHECK(tr_set_mesh(ctx, pos, nrm, uv, 3, idx, 6));
    CHECK(tr_set_primitives(ctx, &prim, 1));
    CHECK(tr_set_instances(ctx, &inst, 1));
    CHECK(tr_set_materials(ctx, &mat, 1));
    CHECK(tr_set_lights(ctx, NULL, 0));
    CHECK(tr_set_uniforms(ctx, &u));
    /* the GGX LUT the transmissive pass samples (the reference loads ggx_lut.png, src/main.rs:295-330); a flat stand-in here */
    uint8_t lut[4 * 4 * 4];
    for (int i = 0; i < 16; i++) { lut[i * 4] = 128; lut[i * 4 + 1] = 32; lut[i * 4 + 2] = 0; lut[i * 4 + 3] = 255; }
    CHECK(tr_set_ggx_lut(ctx, lut, 4, 4));
    tr_write_cluster_data_push_constants wc;
    identity(&wc.inverse_perspective);
    wc.screen_dimensions.x = w; wc.screen_dimensions.y = h;
    CHECK(tr_build_clusters(ctx, &wc));

    tr_frame_params f;
    memset(&f, 0, sizeof(f));
    identity(&f.culling.view);
    f.culling.frustum_x_xz.x = 1.0f; f.culling.frustum_x_xz.y = 1.0f;   /* wide enough never to cull the triangle */
    f.culling.frustum_y_yz.x = 1.0f; f.culling.frustum_y_yz.y = 1.0f;
    f.culling.z_near = 0.0f;
    identity(&f.assign_lights.view_matrix);
    f.assign_lights.view_rotation.w = 1.0f;
    identity(&f.push_constants.proj_view);
    f.push_constants.view_position.z = 5.0f;
    f.push_constants.framebuffer_size.x = w; f.push_constants.framebuffer_size.y = h;
    f.tonemap.a = 1.6f; f.tonemap.d = 0.977f; f.tonemap.b = 1.0f; f.tonemap.c = 1.0f; f.tonemap.saturation = 1.0f;
    CHECK(tr_frame(ctx, &f));

    uint8_t* rgba = (uint8_t*)malloc((size_t)w * h * 4);
    CHECK(tr_read_srgb8(ctx, rgba));
    unsigned long covered = 0;
    for (uint32_t i = 0; i < w * h; i++) covered += (rgba[i * 4] | rgba[i * 4 + 1] | rgba[i * 4 + 2]) != 0;
    printf("frame %ux%u: %lu lit pixels\n", w, h, covered);
    free(rgba);
    CHECK(tr_destroy(ctx));
    return 0;
}

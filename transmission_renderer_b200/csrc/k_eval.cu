// k_eval.cu — batch evaluators that keep the glam-pbr pub fn contracts at the C ABI
// (tr_eval_basic_brdf / tr_eval_transmission_btdf / tr_eval_ibl_volume_refraction).
// Reference: glam-pbr/src/lib.rs:377-423, 200-233, 292-354.  They run the very same
// device functions as the frame kernels (tr_device_pbr.cuh), one element per thread.
#include "tr_internal.h"

using namespace trd;

namespace {

__device__ __forceinline__ f3 ld3(const tr_vec3& v) { return mk3(v.x, v.y, v.z); }
__device__ __forceinline__ MaterialParams ldm(const tr_material_params& m) {
    MaterialParams r;
    r.diffuse_colour = ld3(m.diffuse_colour);
    r.metallic = m.metallic;
    r.perceptual_roughness = m.perceptual_roughness;
    r.index_of_refraction = m.index_of_refraction;
    r.specular_colour = ld3(m.specular_colour);
    r.specular_factor = m.specular_factor;
    return r;
}

__global__ void eval_basic_brdf_kernel(uint32_t n, const tr_basic_brdf_params* in, tr_brdf_result* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    BasicBrdfParams p;
    p.normal = ld3(in[i].normal);
    p.light = ld3(in[i].light);
    p.light_intensity = ld3(in[i].light_intensity);
    p.view = ld3(in[i].view);
    p.material_params = ldm(in[i].material_params);
    BrdfResult r = basic_brdf(p);
    out[i].diffuse.x = r.diffuse.x; out[i].diffuse.y = r.diffuse.y; out[i].diffuse.z = r.diffuse.z;
    out[i].specular.x = r.specular.x; out[i].specular.y = r.specular.y; out[i].specular.z = r.specular.z;
}

__global__ void eval_btdf_kernel(uint32_t n, const tr_transmission_btdf_params* in, tr_vec3* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 r = transmission_btdf(ldm(in[i].material_params), ld3(in[i].normal), ld3(in[i].view), ld3(in[i].light));
    out[i].x = r.x; out[i].y = r.y; out[i].z = r.z;
}

// the frame kernels' light loop for one (pixel, point light) pair: make_pixel_shading -> make_loop_pixel ->
// point_light_lean -> finish_sums, exactly the sequence of shade_kernel (k_shade.cu)
__global__ void eval_point_light_kernel(uint32_t n, const tr_point_light_params* in, tr_point_light_result* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const MaterialParams m = ldm(in[i].material_params);
    const PixelShading ps = make_pixel_shading(m, ld3(in[i].normal), ld3(in[i].view), true);
    const LoopPixel lp = make_loop_pixel(ps, ld3(in[i].position));
    LoopSums sums;
    sums.clear();
    point_light_lean<true>(lp, ld3(in[i].light_position), dup_colour(ld3(in[i].light_colour)), [](const f3&, float&) {}, sums);
    f3 diff, spec, trans;
    finish_sums(sums, ps.f0, ps.df, ps.c_diff_pi, ps.base, lp.a2, lp.at2, true, diff, spec, trans);
    out[i].diffuse.x = diff.x; out[i].diffuse.y = diff.y; out[i].diffuse.z = diff.z;
    out[i].specular.x = spec.x; out[i].specular.y = spec.y; out[i].specular.z = spec.z;
    out[i].transmission.x = trans.x; out[i].transmission.y = trans.y; out[i].transmission.z = trans.z;
}

__global__ void eval_ibl_kernel(uint32_t n, const __grid_constant__ mat4 pv, const tr_ibl_volume_refraction_params* in,
                                tr_vec3* out, const __grid_constant__ PyramidDesc pyr, const __grid_constant__ LutDesc lut) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    IblVolumeRefractionParams p;
    p.material_params = ldm(in[i].material_params);
    p.normal = ld3(in[i].normal);
    p.view = ld3(in[i].view);
    p.position = ld3(in[i].position);
    p.thickness = in[i].thickness;
    p.model_scale = in[i].model_scale;
    p.attenuation_distance = in[i].attenuation_distance;
    p.attenuation_colour = ld3(in[i].attenuation_colour);
    PixelShading s = make_pixel_shading(p.material_params, p.normal, p.view, false);
    f3 r = ibl_volume_refraction(p, pv, log2f((float)in[i].framebuffer_size_x), pyr, lut, s.f0, s.df);
    out[i].x = r.x; out[i].y = r.y; out[i].z = r.z;
}

}  // namespace

namespace tr {
int32_t launch_eval_basic_brdf(uint32_t n, const tr_basic_brdf_params* in, tr_brdf_result* out, cudaStream_t s) {
    if (!n) return TR_OK;
    eval_basic_brdf_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, in, out);
    count_launches(1);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}
int32_t launch_eval_transmission_btdf(uint32_t n, const tr_transmission_btdf_params* in, tr_vec3* out, cudaStream_t s) {
    if (!n) return TR_OK;
    eval_btdf_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, in, out);
    count_launches(1);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}
int32_t launch_eval_point_light(uint32_t n, const tr_point_light_params* in, tr_point_light_result* out, cudaStream_t s) {
    if (!n) return TR_OK;
    eval_point_light_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, in, out);
    count_launches(1);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}
int32_t launch_eval_ibl(uint32_t n, const trd::mat4& pv, const tr_ibl_volume_refraction_params* in, tr_vec3* out,
                        const trd::PyramidDesc& pyr, const trd::LutDesc& lut, cudaStream_t s) {
    if (!n) return TR_OK;
    eval_ibl_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, pv, in, out, pyr, lut);
    count_launches(1);
    TR_CUDA(cudaGetLastError());
    return TR_OK;
}
}  // namespace tr

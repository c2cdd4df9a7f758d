"""include/tr_host.hpp — the C++ host side above the C ABI (the reference's host is compiled code) — against the Python
restatement of the same reference code that the parity tests use (transmission_renderer_b200/host.py, scenes.Camera).
Compiled with g++ here; no GPU and no libtr.so call is involved (the Renderer class is only compiled)."""
import os
import subprocess

import numpy as np
import pytest

from transmission_renderer_b200 import abi, host, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dump(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host_mirror") / "host_mirror_dump")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "host_mirror_dump.cpp"), "-o", exe])
    return exe


def close(a, b, scale=None):
    """Computed floats: a few ulp of the struct's largest value (different summation order / libm); everything else exact."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = max(float(np.abs(b).max()), 1e-30) if scale is None else scale
    return np.abs(a - b).max() <= 4 * np.finfo(np.float32).eps * s


@pytest.mark.parametrize("w,h,pos,yaw,pitch", [(3840, 2160, (0.0, 3.0, 1.0), 0.0, -15.0), (1920, 1080, (0.3, 1.0, -2.0), 30.0, -10.0),
                                                (333, 187, (-12.5, 7.25, 40.0), 140.0, 22.5), (7680, 4320, (0.0, 2.0, 5.0), -75.0, -5.0)])
def test_host_structs_match_the_python_restatement(dump, w, h, pos, yaw, pitch):
    raw = subprocess.run([dump, str(w), str(h), *[repr(float(v)) for v in pos], repr(float(yaw)), repr(float(pitch))],
                         check=True, capture_output=True).stdout
    sizes = [abi.frame_params.itemsize, abi.uniforms.itemsize, abi.write_cluster_data_push_constants.itemsize, abi.light.itemsize,
             abi.light.itemsize, 32]
    assert len(raw) == sum(sizes) and sizes[0] == 304
    off = np.cumsum([0] + sizes)
    part = [raw[off[i]:off[i + 1]] for i in range(len(sizes))]
    cam = scenes.Camera(w, h, pos, yaw, pitch)

    f = np.frombuffer(part[0], abi.frame_params)
    ref = cam.frame_params(host.default_tonemap_params(), flags=abi.TR_FRAME_SKIP_TONEMAP, acceleration_structure_address=0x1234)
    assert int(f["flags"][0]) == abi.TR_FRAME_SKIP_TONEMAP
    pc, rpc = f["push_constants"][0], ref["push_constants"][0]
    assert tuple(pc["framebuffer_size"]) == (w, h) and int(pc["acceleration_structure_address"]) == 0x1234
    assert np.array_equal(pc["view_position"][:3], rpc["view_position"][:3])
    assert close(pc["proj_view"], rpc["proj_view"])
    cu, rcu = f["culling"][0], ref["culling"][0]
    assert close(cu["view"], rcu["view"]) and float(cu["z_near"]) == float(rcu["z_near"])
    assert close(cu["frustum_x_xz"], rcu["frustum_x_xz"], 1.0) and close(cu["frustum_y_yz"], rcu["frustum_y_yz"], 1.0)
    al, ral = f["assign_lights"][0], ref["assign_lights"][0]
    assert close(al["view_matrix"], ral["view_matrix"]) and close(al["view_rotation"], ral["view_rotation"], 1.0)
    tm, rtm = f["tonemap"][0], ref["tonemap"][0]
    for k in ("a", "b", "c", "d", "crosstalk", "saturation", "cross_saturation"):
        assert close(tm[k], rtm[k]), k

    u, ru = np.frombuffer(part[1], abi.uniforms)[0], host.make_uniforms(w, h)[0]
    for k in ("z_near", "z_far", "num_depth_slices", "num_clusters", "cluster_size_in_pixels", "debug_clusters", "ggx_lut_texture_index"):
        assert np.array_equal(u[k], ru[k]), k
    assert close(u["scale"], ru["scale"]) and close(u["bias"], ru["bias"])
    assert close(u["sun_dir"][:3], ru["sun_dir"][:3], 1.0) and np.array_equal(u["sun_intensity"][:3], ru["sun_intensity"][:3])

    wc, rwc = np.frombuffer(part[2], abi.write_cluster_data_push_constants)[0], cam.write_cluster_data()[0]
    assert tuple(wc["screen_dimensions"]) == (w, h) and close(wc["inverse_perspective"], rwc["inverse_perspective"])

    lp = np.frombuffer(part[3], abi.light)
    assert lp.tobytes() == host.light_new_point((0.5, 3.0, 1.5), (1.0, 0.8, 0.6), 8.0).tobytes()
    ls, rls = np.frombuffer(part[4], abi.light)[0], host.light_new_spot((-8.0, 9.0, -14.0), (1.0, 0.9, 0.8), 40.0, (0.0, -1.0, 0.0), 0.3, 0.5)[0]
    for k in ("colour_emission_and_falloff_distance_sq", "spotlight_direction_and_outer_angle"):
        assert np.array_equal(ls[k], rls[k]), k
    assert np.array_equal(ls["position_and_spotlight_epsilon"][:3], rls["position_and_spotlight_epsilon"][:3])
    assert close(ls["position_and_spotlight_epsilon"][3], rls["position_and_spotlight_epsilon"][3], 1.0)

    misc = np.frombuffer(part[5], np.uint32)
    assert misc.tolist() == [host.mip_levels_for_size(w, h), 157, *host.band_rows(h, 3, 8), host.NUM_CLUSTERS, 0,
                             host.mip_levels_for_size(7680, 4320), h]


def test_cpp_host_links_and_fails_loudly_without_a_device(tmp_path):
    """examples/frame.cpp (one frame through tr::Renderer): valid C++17 against include/tr_host.hpp + tr_abi.h, links with
    libtr.so, and without a CUDA device the constructor throws tr::Error with the library's message — no fallback."""
    libdir = os.path.join(ROOT, "transmission_renderer_b200")
    exe = str(tmp_path / "frame_cpp")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "frame.cpp"),
                           "-L", libdir, "-ltr", f"-Wl,-rpath,{libdir}", "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), timeout=120)
    assert p.returncode == 3 and "tr_create" in p.stderr and "no CPU path" in p.stderr

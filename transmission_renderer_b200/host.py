"""Host-side inputs of the path: what the reference computes on the CPU each frame
before recording (`src/main.rs`), restated in float32 numpy.

These are *inputs* of the kernels (matrices, frustum constants, uniforms,
lights); none of this is the hot path.
"""
import math

import numpy as np

from . import abi

f32 = np.float32

Z_NEAR = f32(0.01)   # src/main.rs:56
Z_FAR = f32(500.0)   # src/main.rs:57
NUM_CLUSTERS_X = 24  # src/main.rs:60
NUM_CLUSTERS_Y = 16  # src/main.rs:61
NUM_DEPTH_SLICES = 16  # src/main.rs:62
NUM_CLUSTERS = NUM_CLUSTERS_X * NUM_CLUSTERS_Y * NUM_DEPTH_SLICES


def perspective_matrix_reversed(width, height):
    """src/main.rs:39-54 (math convention: result[row, col])."""
    aspect_ratio = f32(width) / f32(height)
    vertical_fov = f32(math.radians(59.0))
    focal_length = f32(1.0) / f32(math.tan(float(vertical_fov / f32(2.0))))
    a = Z_NEAR / (Z_FAR - Z_NEAR)
    b = Z_FAR * a
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0] = focal_length / aspect_ratio
    m[1, 1] = -focal_length
    m[2, 2] = a
    m[3, 2] = f32(-1.0)
    m[2, 3] = b
    return m


def _normalize(v):
    v = np.asarray(v, dtype=f32)
    return (v * (f32(1.0) / np.sqrt(np.dot(v, v), dtype=f32))).astype(f32)


def look_at_rh(eye, center, up):
    """glam::Mat4::look_at_rh (src/main.rs:520-526)."""
    eye = np.asarray(eye, dtype=f32)
    f = _normalize(np.asarray(center, dtype=f32) - eye)
    s = _normalize(np.cross(f, np.asarray(up, dtype=f32)).astype(f32))
    u = np.cross(s, f).astype(f32)
    m = np.eye(4, dtype=f32)
    m[0, :3] = s
    m[1, :3] = u
    m[2, :3] = -f
    m[0, 3] = -np.dot(s, eye)
    m[1, 3] = -np.dot(u, eye)
    m[2, 3] = np.dot(f, eye)
    return m


def camera_from_yaw_pitch(position, yaw_deg, pitch_deg):
    """dolly YawPitch rig (src/main.rs:514-526): returns (view matrix, rotation quat xyzw)."""
    yaw, pitch = math.radians(yaw_deg), math.radians(pitch_deg)
    # dolly: rotation = from_euler(YXZ, yaw, pitch, 0); forward = rot * -Z, up = rot * Y
    cy, sy = math.cos(yaw / 2), math.sin(yaw / 2)
    cp, sp = math.cos(pitch / 2), math.sin(pitch / 2)
    # q = qy(yaw) * qx(pitch)
    q = np.array([cy * sp, sy * cp, -sy * sp, cy * cp], dtype=f32)  # x y z w
    fwd = quat_rotate(q, np.array([0, 0, -1], dtype=f32))
    up = quat_rotate(q, np.array([0, 1, 0], dtype=f32))
    position = np.asarray(position, dtype=f32)
    return look_at_rh(position, position + fwd, up), q


def quat_rotate(q, v):
    b = q[:3].astype(np.float64)
    w = float(q[3])
    v = np.asarray(v, dtype=np.float64)
    r = v * (w * w - b.dot(b)) + b * (2.0 * v.dot(b)) + np.cross(b, v) * (2.0 * w)
    return r.astype(f32)


def quat_inverse(q):
    return np.array([-q[0], -q[1], -q[2], q[3]], dtype=f32)


def sun_as_normal(pitch=1.1, yaw=4.8):
    """src/main.rs:2715-2722, defaults src/main.rs:531-534."""
    pitch, yaw = f32(pitch), f32(yaw)
    return np.array([np.cos(pitch) * np.sin(yaw), np.sin(pitch), np.cos(pitch) * np.cos(yaw)], dtype=f32)


def light_cluster_coefficients(z_near=Z_NEAR, z_far=Z_FAR, slices=NUM_DEPTH_SLICES):
    """shared-structs/src/lib.rs:44-52."""
    z_near, z_far = f32(z_near), f32(z_far)
    l = np.log2(z_far / z_near, dtype=f32)
    scale = f32(slices) / l
    bias = -(f32(slices) * np.log2(z_near, dtype=f32) / l)
    return z_near, z_far, f32(scale), f32(bias), slices


def make_uniforms(width, height, sun_dir=None, sun_intensity=(3.0, 3.0, 3.0)):
    """src/main.rs:536-552."""
    u = np.zeros(1, dtype=abi.uniforms)
    zn, zf, scale, bias, slices = light_cluster_coefficients()
    u["z_near"], u["z_far"], u["scale"], u["bias"], u["num_depth_slices"] = zn, zf, scale, bias, slices
    sd = sun_as_normal() if sun_dir is None else np.asarray(sun_dir, dtype=f32)
    u["sun_dir"][0, :3] = sd
    u["sun_intensity"][0, :3] = np.asarray(sun_intensity, dtype=f32)
    u["num_clusters"] = (NUM_CLUSTERS_X, NUM_CLUSTERS_Y)
    u["cluster_size_in_pixels"] = (f32(width) / f32(NUM_CLUSTERS_X), f32(height) / f32(NUM_CLUSTERS_Y))
    return u


def make_push_constants(proj_view, view_position, width, height):
    pc = np.zeros(1, dtype=abi.push_constants)
    pc["proj_view"][0] = np.asarray(proj_view, dtype=f32).T
    pc["view_position"][0, :3] = np.asarray(view_position, dtype=f32)
    pc["framebuffer_size"] = (width, height)
    return pc


def make_culling_push_constants(view, perspective):
    """src/main.rs:1728-1746."""
    perspective = np.asarray(perspective, dtype=f32)
    fx = _normalize(perspective[3, :3] + perspective[0, :3])
    fy = _normalize(perspective[3, :3] + perspective[1, :3])
    c = np.zeros(1, dtype=abi.culling_push_constants)
    c["view"][0] = np.asarray(view, dtype=f32).T
    c["frustum_x_xz"] = (fx[0], fx[2])
    c["frustum_y_yz"] = (fy[1], fy[2])
    c["z_near"] = Z_NEAR
    return c


def make_write_cluster_data_push_constants(perspective, width, height):
    """src/main.rs:1502-1505: inverse computed on the host (glam Mat4::inverse)."""
    w = np.zeros(1, dtype=abi.write_cluster_data_push_constants)
    inv = np.linalg.inv(np.asarray(perspective, dtype=np.float64)).astype(f32)
    w["inverse_perspective"][0] = inv.T
    w["screen_dimensions"] = (width, height)
    return w


def make_assign_lights_push_constants(view, camera_rotation):
    """src/main.rs:1785-1788: view_rotation = camera_rotation.inverse()."""
    a = np.zeros(1, dtype=abi.assign_lights_push_constants)
    a["view_matrix"][0] = np.asarray(view, dtype=f32).T
    a["view_rotation"][0] = quat_inverse(np.asarray(camera_rotation, dtype=f32))
    return a


def light_new_point(position, colour, intensity):
    """shared-structs/src/lib.rs:94-103."""
    l = np.zeros(1, dtype=abi.light)
    l["position_and_spotlight_epsilon"][0, :3] = np.asarray(position, dtype=f32)
    l["colour_emission_and_falloff_distance_sq"][0, :3] = np.asarray(colour, dtype=f32) * f32(intensity)
    l["colour_emission_and_falloff_distance_sq"][0, 3] = f32(intensity) / f32(0.05)
    return l


def light_new_spot(position, colour, intensity, direction, inner_angle_rad, outer_angle_rad):
    """shared-structs/src/lib.rs:105-123."""
    l = light_new_point(position, colour, intensity)
    l["position_and_spotlight_epsilon"][0, 3] = np.cos(f32(inner_angle_rad)) - np.cos(f32(outer_angle_rad))
    l["spotlight_direction_and_outer_angle"][0, :3] = np.asarray(direction, dtype=f32)
    l["spotlight_direction_and_outer_angle"][0, 3] = f32(outer_angle_rad)
    return l


def default_tonemap_params():
    """Caller-supplied in the reference (colstodian BakedLottesTonemapperParams::from(default),
    src/main.rs:506-510; colstodian is not in /root/reference).  These follow the published Lottes
    parameterisation with contrast 1.6, shoulder 0.977, hdr_max 8, mid_in 0.18, mid_out 0.267."""
    contrast, shoulder, hdr_max, mid_in, mid_out = 1.6, 0.977, 8.0, 0.18, 0.267
    a = contrast
    d = shoulder
    b = (-(mid_in ** a) + (hdr_max ** a) * mid_out) / ((((hdr_max ** a) ** d) - ((mid_in ** a) ** d)) * mid_out)
    c = ((hdr_max ** a) ** d * (mid_in ** a) - (hdr_max ** a) * ((mid_in ** a) ** d) * mid_out) / (
        (((hdr_max ** a) ** d) - ((mid_in ** a) ** d)) * mid_out)
    p = np.zeros(1, dtype=abi.baked_lottes_tonemapper_params)
    p["a"], p["b"], p["c"], p["d"] = a, b, c, d
    p["crosstalk"], p["saturation"], p["cross_saturation"] = 4.0, contrast, 16.0
    return p


def mip_levels_for_size(width, height):
    """src/main.rs:2590-2592."""
    return int(np.float32(np.log2(np.float32(min(width, height))))) + 1


def band_rows(height, rank, world_size):
    """Image-band partition: rows [floor(r*H/N), floor((r+1)*H/N))."""
    return (rank * height) // world_size, ((rank + 1) * height) // world_size

"""Procedural inputs for the BASELINE.json configs (glTF-Sample-Models is unavailable offline, so
these stand in for src/model_loading.rs).  Everything is generated from a
counter-based hash (splitmix64) so the same seed gives the same bytes everywhere.

These are INPUT generators (host side), not part of the measured path.
"""
import math

import numpy as np

from . import abi, host

f32 = np.float32
U64 = np.uint64


def splitmix64(x):
    x = (np.asarray(x, dtype=U64) + U64(0x9E3779B97F4A7C15))
    with np.errstate(over="ignore"):
        z = x
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        z = z ^ (z >> U64(31))
    return z


def hash01(seed, index):
    """uniform float32 in [0,1) with a 24-bit mantissa, from splitmix64(seed ^ index)."""
    z = splitmix64(U64(seed) ^ np.asarray(index, dtype=U64))
    return ((z >> U64(40)).astype(np.float64) / float(1 << 24)).astype(f32)


# --------------------------------------------------------------------------- meshes
def uv_sphere(segments=32, rings=16):
    """Unit UV sphere, CCW seen from outside (glTF front face)."""
    pos, nrm, uv = [], [], []
    for r in range(rings + 1):
        th = math.pi * r / rings
        for s in range(segments + 1):
            ph = 2 * math.pi * s / segments
            p = (math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph))
            pos.append(p)
            nrm.append(p)
            uv.append((s / segments, r / rings))
    idx = []
    for r in range(rings):
        for s in range(segments):
            a = r * (segments + 1) + s
            b = a + segments + 1
            if r != 0:
                idx += [a, a + 1, b]
            if r != rings - 1:
                idx += [a + 1, b + 1, b]
    return (np.array(pos, f32), np.array(nrm, f32), np.array(uv, f32), np.array(idx, np.uint32))


def box_mesh():
    pos, nrm, uv, idx = [], [], [], []
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    for n, u, v in faces:
        n, u, v = np.array(n, f32), np.array(u, f32), np.array(v, f32)
        base = len(pos)
        for (a, b) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            pos.append(n + a * u + b * v)
            nrm.append(n)
            uv.append(((a + 1) / 2, (b + 1) / 2))
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return np.array(pos, f32), np.array(nrm, f32), np.array(uv, f32), np.array(idx, np.uint32)


def quad_mesh(size=1.0):
    """Ground quad in the xz plane facing +y."""
    pos = np.array([(-size, 0, -size), (-size, 0, size), (size, 0, size), (size, 0, -size)], f32)
    nrm = np.tile(np.array([(0, 1, 0)], f32), (4, 1))
    uv = np.array([(0, 0), (0, 1), (1, 1), (1, 0)], f32)
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    return pos, nrm, uv, idx


def torus_knot(p=2, q=3, n_u=512, n_v=64, tube=0.22, displacement=0.15, seed=0x5EED0003):
    """Displaced torus knot: the "dragon-like" transmissive mesh of config 3."""
    us = np.arange(n_u, dtype=np.float64) / n_u * 2 * math.pi
    def centre(t):
        r = 0.6 + 0.25 * np.cos(q * t)
        return np.stack([r * np.cos(p * t), 0.35 * np.sin(q * t), r * np.sin(p * t)], -1)
    c = centre(us)
    d = centre(us + 1e-4) - c
    T = d / np.linalg.norm(d, axis=1, keepdims=True)
    up = np.array([0.0, 1.0, 0.0])
    N = np.cross(T, up)
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    B = np.cross(T, N)
    vs = np.arange(n_v, dtype=np.float64) / n_v * 2 * math.pi
    # smooth pseudo-fbm displacement from a few hashed harmonics
    amp = np.zeros((n_u, n_v))
    for k in range(1, 5):
        a, b2, ph = hash01(seed, 3 * k), hash01(seed, 3 * k + 1), hash01(seed, 3 * k + 2)
        amp += (0.5 ** k) * np.sin(k * (3 * us[:, None] + 2 * vs[None, :]) * (1 + float(a)) + 6.28 * float(ph)) * (0.5 + float(b2))
    rad = tube * (1.0 + displacement / tube * 0.5 * amp)
    ring = np.cos(vs)[None, :, None] * N[:, None, :] + np.sin(vs)[None, :, None] * B[:, None, :]
    pos = c[:, None, :] + rad[:, :, None] * ring
    pos = pos.reshape(-1, 3)
    idx = []
    for i in range(n_u):
        for j in range(n_v):
            a = i * n_v + j
            b = ((i + 1) % n_u) * n_v + j
            a1 = i * n_v + (j + 1) % n_v
            b1 = ((i + 1) % n_u) * n_v + (j + 1) % n_v
            idx += [a, a1, b, a1, b1, b]
    idx = np.array(idx, np.uint32)
    # smooth vertex normals from the faces
    tri = idx.reshape(-1, 3)
    fn = np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]])
    nrm = np.zeros_like(pos)
    for k in range(3):
        np.add.at(nrm, tri[:, k], fn)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    # orient outward (away from the tube centre line)
    out = pos - np.repeat(c, n_v, axis=0)
    flip = np.sum(nrm * out, axis=1) < 0
    if flip.mean() > 0.5:
        nrm = -nrm
        idx = idx.reshape(-1, 3)[:, [0, 2, 1]].reshape(-1)
    uv = np.stack([np.repeat(np.arange(n_u) / n_u, n_v), np.tile(np.arange(n_v) / n_v, n_u)], -1)
    return pos.astype(f32), nrm.astype(f32), uv.astype(f32), idx.astype(np.uint32)


class MeshSet:
    """Concatenated vertex/index arrays + one PrimitiveInfo per mesh (ModelStagingBuffers, main.rs:2495-2560)."""

    def __init__(self):
        self.pos, self.nrm, self.uv, self.idx, self.prims = [], [], [], [], []
        self.n_vertices = 0
        self.n_indices = 0

    def add(self, mesh, draw_buffer_index):
        pos, nrm, uv, idx = mesh
        lo, hi = pos.min(axis=0), pos.max(axis=0)
        centre = ((lo + hi) / 2).astype(f32)                         # model_loading.rs:148-155
        radius = f32(np.linalg.norm((hi - lo).astype(np.float64)) / 2)
        p = np.zeros(1, dtype=abi.primitive_info)
        p["packed_bounding_sphere"][0] = (*centre, radius)
        p["draw_buffer_index"] = draw_buffer_index
        p["index_count"] = len(idx)
        p["first_index"] = self.n_indices
        p["first_instance"] = len(self.prims)
        self.pos.append(pos)
        self.nrm.append(nrm)
        self.uv.append(uv)
        self.idx.append(idx + np.uint32(self.n_vertices))
        self.prims.append(p)
        self.n_vertices += len(pos)
        self.n_indices += len(idx)
        return len(self.prims) - 1

    def arrays(self):
        return dict(positions=np.concatenate(self.pos), normals=np.concatenate(self.nrm), uvs=np.concatenate(self.uv),
                    indices=np.concatenate(self.idx)), np.concatenate(self.prims)


def make_instance(translation, scale, rotation, primitive_id, material_id):
    i = np.zeros(1, dtype=abi.instance)
    i["translation_and_scale"][0] = (*translation, scale)
    i["rotation"][0] = rotation
    i["primitive_id"] = primitive_id
    i["material_id"] = material_id
    return i


# --------------------------------------------------------------------------- cameras
class Camera:
    def __init__(self, width, height, position=(0.0, 3.0, 1.0), yaw_deg=0.0, pitch_deg=-15.0):
        self.width, self.height = width, height
        self.position = np.asarray(position, f32)
        self.view, self.rotation = host.camera_from_yaw_pitch(position, yaw_deg, pitch_deg)  # main.rs:514-526
        self.perspective = host.perspective_matrix_reversed(width, height)
        self.proj_view = (self.perspective.astype(f32) @ self.view.astype(f32)).astype(f32)   # main.rs:1188-1195

    def push_constants(self, acceleration_structure_address=0):
        pc = host.make_push_constants(self.proj_view, self.position, self.width, self.height)
        pc["acceleration_structure_address"] = acceleration_structure_address
        return pc

    def culling(self):
        return host.make_culling_push_constants(self.view, self.perspective)

    def write_cluster_data(self):
        return host.make_write_cluster_data_push_constants(self.perspective, self.width, self.height)

    def assign_lights(self):
        return host.make_assign_lights_push_constants(self.view, self.rotation)

    def frame_params(self, tonemap=None, flags=0, acceleration_structure_address=0):
        f = np.zeros(1, dtype=abi.frame_params)
        f["culling"] = self.culling()
        f["assign_lights"] = self.assign_lights()
        f["push_constants"] = self.push_constants(acceleration_structure_address)
        f["tonemap"] = host.default_tonemap_params() if tonemap is None else tonemap
        f["flags"] = flags
        return f


# --------------------------------------------------------------------------- config 1 (synthetic G-buffer)
def procedural_opaque_frame(width, height, seed=0x5EED0001, checker=32):
    """rgb = 4 * hash01(x, y, c) * checker(32 px), alpha 1 -> float32 (h, w, 4)."""
    y, x = np.mgrid[0:height, 0:width]
    img = np.ones((height, width, 4), dtype=f32)
    chk = (((x // checker) + (y // checker)) & 1).astype(f32)
    for c in range(3):
        idx = (y.astype(np.uint64) * np.uint64(width) + x.astype(np.uint64)) * np.uint64(3) + np.uint64(c)
        img[..., c] = f32(4.0) * hash01(seed, idx) * (f32(0.25) + f32(0.75) * chk)
    return img


def config1(size=512, roughness=0.25, lights=()):
    """BASELINE configs[0]: BRDF + transmission_btdf + ibl_volume_refraction over a synthetic size^2 G-buffer,
    one directional light (the sun), roughness 0.25 (SURVEY.md 8d "Config 1")."""
    w = h = size
    cam = Camera(w, h, (0.0, 3.0, 1.0), 0.0, -15.0)
    y, x = np.mgrid[0:h, 0:w]
    sx = ((x + 0.5) / w * 2 - 1).astype(f32)
    sy = ((y + 0.5) / h * 2 - 1).astype(f32)
    r2 = sx * sx + sy * sy
    inside = r2 < 1.0
    # keep the normals off the exact silhouette (n.v -> 0 makes the BTDF's visibility term singular)
    nz = np.maximum(np.sqrt(np.maximum(1.0 - r2, 0.0)), 0.2).astype(f32)
    # camera-facing basis so the disc is a sphere seen from the camera
    right = cam.view[0, :3]
    up = cam.view[1, :3]
    back = cam.view[2, :3]
    n_local = np.where(inside[..., None], np.stack([sx, -sy, nz], -1), np.array([0, 0, 1], f32))
    n_local = n_local / np.linalg.norm(n_local, axis=-1, keepdims=True)
    normal = (n_local[..., 0:1] * right + n_local[..., 1:2] * up + n_local[..., 2:3] * back).astype(f32)
    centre = np.array([0.0, 2.0, 0.0], f32)
    position = (centre + normal).astype(f32)
    # depth of the projected position (frag_coord.z), reversed-Z
    ph = np.concatenate([position, np.ones((h, w, 1), f32)], -1) @ cam.proj_view.T
    depth = (ph[..., 2] / ph[..., 3]).astype(f32)
    assert depth.min() > 0 and depth.max() < 1
    material = abi.default_material(1)
    material["diffuse_factor"] = (0.8, 0.8, 0.8, 1.0)
    material["metallic_factor"] = 0.0
    material["roughness_factor"] = roughness
    material["index_of_refraction"] = 1.5
    material["transmission_factor"] = 1.0
    material["thickness_factor"] = 1.0
    material["attenuation_distance"] = 1.0
    material["attenuation_colour"] = (0.9, 0.4, 0.2, 0.0)
    gbuffer = dict(depth=depth, normal=normal, uv=np.zeros((h, w, 2), f32), material_id=np.zeros((h, w), np.uint32),
                   scale=np.ones((h, w), f32), position=position)
    lights = np.concatenate(lights) if len(lights) else np.zeros(0, dtype=abi.light)
    return dict(camera=cam, gbuffer=gbuffer, materials=material, lights=lights, uniforms=host.make_uniforms(w, h),
                opaque=procedural_opaque_frame(w, h))


# --------------------------------------------------------------------------- analytic sphere G-buffer (no rasteriser needed)
def raycast_spheres(cam, centres, radii, material_ids, scale_plane=False):
    """Exact ray/sphere G-buffer (nearest hit) — used to exercise K4/K6 independently of K3."""
    w, h = cam.width, cam.height
    inv = np.linalg.inv(cam.proj_view.astype(np.float64))
    y, x = np.mgrid[0:h, 0:w]
    ndc = np.stack([(x + 0.5) / w * 2 - 1, (y + 0.5) / h * 2 - 1, np.full((h, w), 0.5), np.ones((h, w))], -1)
    pw = ndc @ inv.T
    pw = pw[..., :3] / pw[..., 3:4]
    o = cam.position.astype(np.float64)
    d = pw - o
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    best_t = np.full((h, w), np.inf)
    best = np.full((h, w), -1, dtype=np.int64)
    for i, (c, r) in enumerate(zip(np.asarray(centres, np.float64), radii)):
        oc = o - c
        b = d @ oc
        disc = b * b - (oc @ oc - r * r)
        t = -b - np.sqrt(np.maximum(disc, 0))
        hit = (disc > 0) & (t > 0.02) & (t < best_t)
        best_t = np.where(hit, t, best_t)
        best = np.where(hit, i, best)
    covered = best >= 0
    pos = o + d * np.where(covered, best_t, 0)[..., None]
    cidx = np.maximum(best, 0)
    nrm = (pos - np.asarray(centres, np.float64)[cidx]) / np.asarray(radii, np.float64)[cidx][..., None]
    ph = np.concatenate([pos, np.ones((h, w, 1))], -1) @ cam.proj_view.astype(np.float64).T
    depth = np.where(covered, ph[..., 2] / ph[..., 3], 0.0).astype(f32)
    g = dict(depth=depth, normal=np.where(covered[..., None], nrm, 0).astype(f32), uv=np.zeros((h, w, 2), f32),
             material_id=np.where(covered, np.asarray(material_ids, np.int64)[cidx], 0xFFFFFFFF).astype(np.uint32),
             scale=np.ones((h, w), f32) if scale_plane else None, position=None)
    return g


def hashed_materials(n, seed, transmissive=False, roughness_range=(0.1, 0.9)):
    m = abi.default_material(n)
    i = np.arange(n, dtype=np.uint64) * np.uint64(8)
    lo, hi = roughness_range
    m["roughness_factor"] = f32(lo) + f32(hi - lo) * hash01(seed, i)
    m["metallic_factor"] = (hash01(seed, i + np.uint64(1)) > 0.5).astype(f32) * (0.0 if transmissive else 1.0)
    m["diffuse_factor"][:, 0] = f32(0.2) + f32(0.8) * hash01(seed, i + np.uint64(2))
    m["diffuse_factor"][:, 1] = f32(0.2) + f32(0.8) * hash01(seed, i + np.uint64(3))
    m["diffuse_factor"][:, 2] = f32(0.2) + f32(0.8) * hash01(seed, i + np.uint64(4))
    if transmissive:
        m["transmission_factor"] = 1.0
        m["thickness_factor"] = f32(0.2) + f32(0.6) * hash01(seed, i + np.uint64(5))
        m["attenuation_distance"] = f32(0.3) + f32(1.0) * hash01(seed, i + np.uint64(6))
        m["attenuation_colour"][:, 0] = f32(0.3) + f32(0.7) * hash01(seed, i + np.uint64(7))
        m["attenuation_colour"][:, 1] = f32(0.3) + f32(0.7) * hash01(seed ^ 0xABC, i)
        m["attenuation_colour"][:, 2] = f32(0.3) + f32(0.7) * hash01(seed ^ 0xABC, i + np.uint64(1))
    return m


def config2_lights():
    """The reference's two point lights (main.rs:450-453) + two more (SURVEY.md 8d "Config 2")."""
    return np.concatenate([
        host.light_new_point((0.0, 0.8, 0.0), (1, 0, 0), 5.0),
        host.light_new_point((8.0, 0.8, 0.0), (0, 1, 0), 10.0),
        host.light_new_point((-4.0, 2.0, 2.0), (1, 1, 1), 20.0),
        host.light_new_point((4.0, 3.0, -2.0), (0.2, 0.3, 1.0), 20.0),
    ])


def hashed_point_lights(n, seed, box=((-30, 0.5, -30), (30, 10, 30)), intensity=(5.0, 50.0)):
    ls = []
    lo, hi = np.asarray(box[0], f32), np.asarray(box[1], f32)
    for i in range(n):
        u = hash01(seed, np.arange(8, dtype=np.uint64) + np.uint64(16 * i))
        pos = lo + (hi - lo) * u[:3]
        col = f32(0.3) + f32(0.7) * u[3:6]
        inten = f32(intensity[0]) + f32(intensity[1] - intensity[0]) * u[6]
        ls.append(host.light_new_point(pos, col, float(inten)))
    return np.concatenate(ls) if ls else np.zeros(0, dtype=abi.light)


# --------------------------------------------------------------------------- full scenes for the rasterised path
def sphere_grid_scene(width, height, seed=0x5EED0002, grid=8, radius=0.4, transmissive_knot=False, ground=True):
    """Configs 2/3: grid x grid UV spheres in front of the camera (+ ground quad, + displaced torus knot)."""
    cam = Camera(width, height, (0.0, 3.0, 6.0), 0.0, -15.0)
    meshes = MeshSet()
    sphere = meshes.add(uv_sphere(32, 16), 0)
    quad = meshes.add(quad_mesh(1.0), 0) if ground else None
    knot = meshes.add(torus_knot(n_u=512 if transmissive_knot else 8, n_v=64 if transmissive_knot else 4), 2) if transmissive_knot else None
    n_s = grid * grid
    mats = [hashed_materials(n_s, seed)]
    inst = []
    for j in range(grid):
        for i in range(grid):
            k = j * grid + i
            pos = ((i - (grid - 1) / 2) * 1.1, 0.5 + (j % 3) * 0.9 + 0.0, -(j * 1.1))
            inst.append(make_instance(pos, radius, (0, 0, 0, 1), sphere, k))
    n_mat = n_s
    if ground:
        g = abi.default_material(1)
        g["diffuse_factor"] = (0.5, 0.5, 0.55, 1.0)
        g["metallic_factor"] = 0.0
        g["roughness_factor"] = 0.7
        mats.append(g)
        inst.append(make_instance((0.0, 0.0, -4.0), 30.0, (0, 0, 0, 1), quad, n_mat))
        n_mat += 1
    if transmissive_knot:
        t = abi.default_material(1)
        t["diffuse_factor"] = (1.0, 1.0, 1.0, 1.0)
        t["metallic_factor"] = 0.0
        t["roughness_factor"] = 0.25
        t["transmission_factor"] = 1.0
        t["thickness_factor"] = 0.5
        t["attenuation_distance"] = 0.5
        t["attenuation_colour"] = (0.9, 0.4, 0.2, 0.0)
        mats.append(t)
        inst.append(make_instance((0.0, 2.2, 2.2), 1.6, (0.0, 0.0, 0.0, 1.0), knot, n_mat))
        n_mat += 1
    mesh, prims = meshes.arrays()
    return dict(camera=cam, mesh=mesh, primitives=prims, instances=np.concatenate(inst), materials=np.concatenate(mats),
                lights=config2_lights(), uniforms=host.make_uniforms(width, height))


def instanced_scene(width, height, n_instances=10000, n_lights=64, seed=0x5EED0004, transmissive_fraction=0.3,
                    yaw_deg=0.0, layout="shells"):
    """Config 4/5: n instances of 8 primitive types in a 60x20x60 box around the camera, 30 % frosted glass,
    hashed point lights with intensity in [5, 50] (falloff radius 10-32 m).

    layout="box"    positions stay uniform in the box (an instance next to the camera then hides everything).
    layout="shells" positions are remapped radially about the camera: frosted glass into the 5-25 m shell,
                    opaque instances beyond 10 m — both layers then cover ~96 % of the frame, the "2 full
                    screens of fragments" worst case the reference's readme names (readme.md:74)."""
    cam = Camera(width, height, (0.0, 6.0, 0.0), yaw_deg, -10.0)
    meshes = MeshSet()
    base = [uv_sphere(24, 12), box_mesh(), uv_sphere(16, 8), torus_knot(n_u=96, n_v=12, seed=seed)]
    prim_opaque = [meshes.add(m, 0) for m in base]
    prim_trans = [meshes.add(m, 2) for m in base]
    n_mat_o, n_mat_t = 64, 32
    mats = np.concatenate([hashed_materials(n_mat_o, seed), hashed_materials(n_mat_t, seed ^ 0x77, True, (0.15, 0.5))])
    k = np.arange(n_instances, dtype=np.uint64) * np.uint64(16)
    u = [hash01(seed, k + np.uint64(j)) for j in range(12)]
    inst = np.zeros(n_instances, dtype=abi.instance)
    inst["translation_and_scale"][:, 0] = (u[0] - f32(0.5)) * f32(60.0)
    inst["translation_and_scale"][:, 1] = u[1] * f32(20.0) - f32(2.0)
    inst["translation_and_scale"][:, 2] = (u[2] - f32(0.5)) * f32(60.0)
    inst["translation_and_scale"][:, 3] = f32(0.25) + f32(1.75) * u[3]
    is_t = u[8] < f32(transmissive_fraction)
    if layout == "shells":
        c = np.asarray(cam.position, np.float64)
        d = inst["translation_and_scale"][:, :3].astype(np.float64) - c
        r = np.linalg.norm(d, axis=1)
        r_max = float(np.sqrt(30.0 ** 2 + 14.0 ** 2 + 30.0 ** 2))
        new_r = np.where(is_t, 5.0 + 20.0 * (r / r_max), r * (1.0 - 10.0 / r_max) + 10.0)
        inst["translation_and_scale"][:, :3] = (c + d / np.maximum(r, 1e-6)[:, None] * new_r[:, None]).astype(f32)
    elif layout != "box":
        raise ValueError(layout)
    q = np.stack([u[4] - f32(0.5), u[5] - f32(0.5), u[6] - f32(0.5), u[7] - f32(0.5)], -1).astype(np.float64)
    q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-9)
    inst["rotation"] = q.astype(f32)
    shape = np.minimum((u[9] * f32(4)).astype(np.int64), 3)
    inst["primitive_id"] = np.where(is_t, np.asarray(prim_trans)[shape], np.asarray(prim_opaque)[shape])
    inst["material_id"] = np.where(is_t, n_mat_o + np.minimum((u[10] * f32(n_mat_t)).astype(np.int64), n_mat_t - 1),
                                   np.minimum((u[10] * f32(n_mat_o)).astype(np.int64), n_mat_o - 1))
    mesh, prims = meshes.arrays()
    lights = hashed_point_lights(n_lights, seed ^ 0x11, box=((-30, 0.5, -30), (30, 12, 30)))
    return dict(camera=cam, mesh=mesh, primitives=prims, instances=inst, materials=mats, lights=lights,
                uniforms=host.make_uniforms(width, height))


# --------------------------------------------------------------------------- procedural material textures (row N2)
def srgb_decode(c8):
    x = np.asarray(c8, np.float64) / 255.0
    return np.where(x <= 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)


def srgb_encode(lin):
    lin = np.clip(np.asarray(lin, np.float64), 0.0, 1.0)
    s = np.where(lin <= 0.0031308, lin * 12.92, 1.055 * lin ** (1 / 2.4) - 0.055)
    return np.floor(s * 255.0 + 0.5).astype(np.uint8)


def make_mips(rgba8, srgb):
    """Full mip chain (src/model_loading.rs:354: levels = floor(log2(min(w, h))) + 1) by 2x2 box filtering — what the
    loader's blit chain produces for power-of-two images; colour of sRGB images is filtered in linear light."""
    levels = [np.ascontiguousarray(rgba8, np.uint8)]
    n = host.mip_levels_for_size(rgba8.shape[1], rgba8.shape[0])
    for _ in range(1, n):
        src = levels[-1]
        h, w = max(1, src.shape[0] // 2), max(1, src.shape[1] // 2)
        lin = src.astype(np.float64) / 255.0
        if srgb:
            lin[..., :3] = srgb_decode(src[..., :3])
        lin = lin[: h * 2, : w * 2] if src.shape[0] >= 2 and src.shape[1] >= 2 else lin
        if src.shape[0] >= 2 and src.shape[1] >= 2:
            lin = lin.reshape(h, 2, w, 2, 4).mean(axis=(1, 3))
        else:
            lin = lin[:h, :w]
        out = np.floor(np.clip(lin, 0, 1) * 255.0 + 0.5).astype(np.uint8)
        if srgb:
            out[..., :3] = srgb_encode(lin[..., :3])
        levels.append(out)
    return levels


def _value_noise(size, cells, seed):
    g = hash01(seed, np.arange(cells * cells, dtype=np.uint64)).reshape(cells, cells).astype(np.float64)
    t = (np.arange(size) + 0.5) / size * cells
    i0 = np.floor(t).astype(int) % cells
    i1 = (i0 + 1) % cells
    f = t - np.floor(t)
    f = f * f * (3 - 2 * f)
    top = g[i0][:, i0] * (1 - f)[None, :] + g[i0][:, i1] * f[None, :]
    bot = g[i1][:, i0] * (1 - f)[None, :] + g[i1][:, i1] * f[None, :]
    return top * (1 - f)[:, None] + bot * f[:, None]


def procedural_textures(size=256, seed=0x5EED00A2):
    """Tileable RGBA8 textures for every slot of `Textures` the shaders read (shared-structs lib.rs:143-153):
    0 diffuse (sRGB, checker + noise), 1 metallic-roughness (UNORM: G roughness, B metallic), 2 normal map (UNORM),
    3 emissive (sRGB, sparse dots), 4 transmission (R), 5 thickness (G), 6 specular (A), 7 specular colour (sRGB)."""
    y, x = np.mgrid[0:size, 0:size]
    n1 = _value_noise(size, 8, seed)
    n2 = _value_noise(size, 16, seed + 1)
    n3 = _value_noise(size, 4, seed + 2)
    chk = (((x // (size // 8)) + (y // (size // 8))) & 1).astype(np.float64)

    def pack(r, g, b, a=1.0):
        img = np.stack(np.broadcast_arrays(r, g, b, a), -1)
        return np.floor(np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)

    diffuse = pack(0.25 + 0.6 * chk * n1 + 0.1 * n2, 0.3 + 0.5 * (1 - chk) * n2, 0.35 + 0.4 * n3, 0.4 + 0.6 * n2)
    metal_rough = pack(0.0 * n1, 0.25 + 0.7 * n1, (n3 > 0.55).astype(np.float64))
    height = 0.6 * n1 + 0.4 * n2
    dx = (np.roll(height, -1, axis=1) - np.roll(height, 1, axis=1)) * size / 16.0
    dy = (np.roll(height, -1, axis=0) - np.roll(height, 1, axis=0)) * size / 16.0
    nrm = np.stack([-dx, -dy, np.ones_like(dx)], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    normal_map = pack(nrm[..., 0] * 0.5 + 0.5, nrm[..., 1] * 0.5 + 0.5, nrm[..., 2] * 0.5 + 0.5)
    dots = ((n2 > 0.8) & (chk > 0)).astype(np.float64)
    emissive = pack(dots, 0.6 * dots, 0.2 * dots)
    transmission = pack(0.35 + 0.65 * n1, 0 * n1, 0 * n1)
    thickness = pack(0 * n1, 0.3 + 0.7 * n3, 0 * n1)
    specular = pack(0 * n1, 0 * n1, 0 * n1, 0.3 + 0.7 * n2)
    specular_colour = pack(0.6 + 0.4 * n1, 0.6 + 0.4 * n2, 0.6 + 0.4 * n3)
    flags = [True, False, False, True, False, False, False, True]
    imgs = [diffuse, metal_rough, normal_map, emissive, transmission, thickness, specular, specular_colour]
    return [dict(levels=make_mips(im, srgb), srgb=srgb) for im, srgb in zip(imgs, flags)]


# indices into abi.material_info["textures"]: diffuse, metallic_roughness, normal_map, emissive, occlusion, transmission,
# thickness, specular, specular_colour
TEX_SLOTS = dict(diffuse=0, metallic_roughness=1, normal_map=2, emissive=3, occlusion=4, transmission=5, thickness=6,
                 specular=7, specular_colour=8)


def textured_sphere_scene(width, height, seed=0x5EED00A3, grid=5, uv_repeat=3.0):
    """Row N2: the config-2/3 layout with texture-mapped materials — every sphere binds a different subset of the
    texture slots (normal maps on half of them), the ground binds diffuse + normal map with uv repeat, and the
    transmissive knot binds transmission + thickness + diffuse."""
    s = sphere_grid_scene(width, height, seed=seed, grid=grid, transmissive_knot=True)
    textures = procedural_textures()
    mats = s["materials"]
    n_s = grid * grid
    bind = [("diffuse", 0), ("metallic_roughness", 1), ("normal_map", 2), ("emissive", 3), ("specular", 6),
            ("specular_colour", 7)]
    for k in range(n_s):
        for j, (slot, tex) in enumerate(bind):
            if (k >> j) & 1 or (slot == "diffuse" and k % 3 == 0):
                mats["textures"][k, TEX_SLOTS[slot]] = tex
        if mats["textures"][k, TEX_SLOTS["emissive"]] != -1:
            mats["emissive_factor"][k, :3] = (2.0, 2.0, 2.0)
    ground, knot = n_s, n_s + 1
    mats["textures"][ground, TEX_SLOTS["diffuse"]] = 0
    mats["textures"][ground, TEX_SLOTS["normal_map"]] = 2
    mats["textures"][ground, TEX_SLOTS["metallic_roughness"]] = 1
    mats["textures"][knot, TEX_SLOTS["diffuse"]] = 0
    mats["textures"][knot, TEX_SLOTS["transmission"]] = 4
    mats["textures"][knot, TEX_SLOTS["thickness"]] = 5
    mats["textures"][knot, TEX_SLOTS["normal_map"]] = 2
    s["mesh"]["uvs"] = (s["mesh"]["uvs"] * f32(uv_repeat)).astype(f32)   # exercises the repeat wrap
    s["textures"] = textures
    return s


def alpha_clip_scene(width, height, seed=0x5EED00A4):
    """Row N3: draw buffers 1 (alpha clip) and 3 (transmission + alpha clip), src/model_loading.rs:68-78: perforated
    spheres in front of solid ones and a perforated glass knot; the holes come from the diffuse texture's alpha
    against `alpha_clipping_cutoff` (depth_pre_pass_alpha_clip, shader/src/lib.rs:269-293)."""
    cam = Camera(width, height, (0.0, 3.0, 6.0), 0.0, -15.0)
    meshes = MeshSet()
    sphere = meshes.add(uv_sphere(32, 16), 0)
    sphere_clip = meshes.add(uv_sphere(32, 16), 1)
    quad = meshes.add(quad_mesh(1.0), 0)
    knot_clip = meshes.add(torus_knot(n_u=256, n_v=32), 3)
    knot = meshes.add(torus_knot(n_u=128, n_v=16), 2)
    textures = procedural_textures()
    grid = 5
    n_s = grid * grid
    mats = hashed_materials(n_s + 3, seed)
    inst = []
    for j in range(grid):
        for i in range(grid):
            k = j * grid + i
            clip = (i + j) % 2 == 0
            pos = ((i - (grid - 1) / 2) * 1.3, 0.6 + (j % 3) * 0.9, -(j * 1.2))
            inst.append(make_instance(pos, 0.55, (0, 0, 0, 1), sphere_clip if clip else sphere, k))
            if clip:
                mats["textures"][k, TEX_SLOTS["diffuse"]] = 0
                mats["alpha_clipping_cutoff"][k] = 0.55 + 0.1 * (k % 3)
                mats["textures"][k, TEX_SLOTS["normal_map"]] = 2 if k % 4 == 0 else -1
    ground, glass_clip, glass = n_s, n_s + 1, n_s + 2
    mats["metallic_factor"][ground] = 0.0
    mats["roughness_factor"][ground] = 0.7
    mats["diffuse_factor"][ground] = (0.5, 0.5, 0.55, 1.0)
    inst.append(make_instance((0.0, 0.0, -4.0), 30.0, (0, 0, 0, 1), quad, ground))
    for m_id, alpha_factor in ((glass_clip, 1.0), (glass, 0.9)):
        mats["metallic_factor"][m_id] = 0.0
        mats["roughness_factor"][m_id] = 0.3
        mats["diffuse_factor"][m_id] = (1.0, 1.0, 1.0, alpha_factor)
        mats["transmission_factor"][m_id] = 1.0
        mats["thickness_factor"][m_id] = 0.5
        mats["attenuation_distance"][m_id] = 0.6
        mats["attenuation_colour"][m_id] = (0.9, 0.5, 0.3, 0.0)
    mats["textures"][glass_clip, TEX_SLOTS["diffuse"]] = 0
    mats["alpha_clipping_cutoff"][glass_clip] = 0.6
    inst.append(make_instance((-1.2, 2.2, 2.4), 1.3, (0.0, 0.0, 0.0, 1.0), knot_clip, glass_clip))
    inst.append(make_instance((1.6, 2.0, 2.0), 1.1, (0.0, 0.38268343, 0.0, 0.92387953), knot, glass))
    mesh, prims = meshes.arrays()
    mesh["uvs"] = (mesh["uvs"] * f32(2.0)).astype(f32)
    return dict(camera=cam, mesh=mesh, primitives=prims, instances=np.concatenate(inst), materials=mats,
                lights=config2_lights(), uniforms=host.make_uniforms(width, height), textures=textures)


# --------------------------------------------------------------------------- ray-queried shadows (`--ray-tracing`)
def shadow_scene(width, height, seed=0x5EED00A5):
    """Occluders over a ground quad, lit by the sun, point lights and a spotlight placed so that shadows fall inside the
    view: spheres and rotated boxes (draw buffer 0), an alpha-clip sphere (buffer 1: casts, src/main.rs:617-620), and a
    frosted-glass knot (buffer 2: receives shadows but casts none)."""
    cam = Camera(width, height, (0.0, 3.2, 6.5), 0.0, -20.0)
    meshes = MeshSet()
    sphere = meshes.add(uv_sphere(32, 16), 0)
    box = meshes.add(box_mesh(), 0)
    quad = meshes.add(quad_mesh(1.0), 0)
    sphere_clip = meshes.add(uv_sphere(24, 12), 1)
    knot = meshes.add(torus_knot(n_u=192, n_v=24), 2)
    mats = hashed_materials(8, seed)
    mats["metallic_factor"][:6] = 0.0
    mats["diffuse_factor"][6] = (0.55, 0.55, 0.6, 1.0)   # ground
    mats["roughness_factor"][6] = 0.8
    mats["metallic_factor"][6] = 0.0
    glass = 7
    mats["diffuse_factor"][glass] = (1.0, 1.0, 1.0, 1.0)
    mats["metallic_factor"][glass] = 0.0
    mats["roughness_factor"][glass] = 0.3
    mats["transmission_factor"][glass] = 1.0
    mats["thickness_factor"][glass] = 0.5
    mats["attenuation_distance"][glass] = 0.8
    mats["attenuation_colour"][glass] = (0.8, 0.9, 0.5, 0.0)
    s45 = (0.0, 0.38268343, 0.0, 0.92387953)
    tilt = (0.25881905, 0.0, 0.0, 0.96592583)
    inst = [
        make_instance((0.0, 0.0, -3.0), 25.0, (0, 0, 0, 1), quad, 6),
        make_instance((-2.2, 0.9, -1.0), 0.9, (0, 0, 0, 1), sphere, 0),
        make_instance((1.8, 0.6, 0.4), 0.6, s45, box, 1),
        make_instance((0.2, 1.9, -2.6), 0.7, tilt, box, 2),
        make_instance((3.4, 1.2, -2.0), 0.5, (0, 0, 0, 1), sphere, 3),
        make_instance((-0.6, 0.45, 1.6), 0.45, (0, 0, 0, 1), sphere_clip, 4),
        make_instance((-3.6, 2.4, -3.4), 0.35, s45, box, 5),
        make_instance((0.6, 1.3, 2.4), 0.9, (0.0, 0.0, 0.0, 1.0), knot, glass),
    ]
    lights = np.concatenate([
        host.light_new_point((0.0, 4.5, 1.0), (1.0, 0.95, 0.9), 40.0),
        host.light_new_point((-4.0, 1.2, 1.5), (1.0, 0.3, 0.2), 25.0),
        host.light_new_point((4.5, 2.0, 1.0), (0.2, 0.4, 1.0), 30.0),
        host.light_new_point((0.5, 0.4, -1.2), (0.3, 1.0, 0.4), 8.0),
        host.light_new_spot((2.5, 5.0, 2.5), (1.0, 1.0, 0.8), 80.0, _unit((-0.35, -0.85, -0.4)), 0.35, 0.55),
        host.light_new_point((-1.5, 3.0, -5.0), (0.9, 0.7, 1.0), 35.0),
    ])
    mesh, prims = meshes.arrays()
    return dict(camera=cam, mesh=mesh, primitives=prims, instances=np.concatenate(inst), materials=mats,
                lights=lights, uniforms=host.make_uniforms(width, height))


def _unit(v):
    v = np.asarray(v, np.float64)
    return tuple((v / np.linalg.norm(v)).astype(f32))

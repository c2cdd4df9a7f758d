"""Design rules of the visibility pass (K3) as statements that can be checked on the CPU, without the kernels: the chunk test of
the binning pass (below) and the band filter's instance test (further down), both against the oracle's rasteriser, and the
error bound of the tile pass's fp32 coarse edge test (last).  The kernels themselves are pinned bit for bit on the GPU
(tests/test_gpu_parity.py); these tests pin the REASONS they may skip work.

The binning pass's chunk test (csrc/k_visibility.cu bin_count_kernel + csrc/tr_api.cu ensure_chunks): a chunk of 64 consecutive triangles whose bounding sphere lies beyond one of the four planes
(frame left / right, two rows above / below the band) by more than the kernel's slack contributes no pixel of the band — so
dropping those triangles leaves every G-buffer plane of the band byte-identical.  The test restates the chunk construction and
the plane test in numpy (the kernel's formulas and margins, float64 arithmetic), removes the chunks it rejects from the index
buffer and lets the ORACLE's rasteriser render the band with and without them, over random cameras, bands and instances — large
and tiny, partly behind the camera, straddling the frame's sides and the band's rows.  (One instance per primitive, as the
reference's loader creates them, so that removing triangles from a primitive removes them for exactly one instance.)"""
import numpy as np
import pytest

from transmission_renderer_b200 import scenes

CHUNK = 64


def chunk_spheres(pos, idx):
    """ensure_chunks: centre of the chunk's box, radius to its farthest vertex (+ the rounding of the stored centre)."""
    tris = idx.reshape(-1, 3)
    out = []
    for t0 in range(0, len(tris), CHUNK):
        v = pos[tris[t0:t0 + CHUNK].reshape(-1)].astype(np.float64)
        c = 0.5 * (v.min(0) + v.max(0))
        fc = c.astype(np.float32)
        r = (np.sqrt(((v - c) ** 2).sum(1).max()) + np.abs(fc - c).sum()) * (1.0 + 1e-6)
        out.append((fc.astype(np.float64), float(np.float32(r))))
    return out


def rotate(q, v):
    b, w = q[:3], q[3]
    return v * (w * w - b @ b) + b * (2.0 * (v @ b)) + np.cross(b, v) * (2.0 * w)


def band_planes(proj_view, height, y0, y1):
    """launch_visibility: clip-space half spaces x >= -w, x <= w and the band's rows with two pixels of slack, as world planes."""
    m = proj_view.astype(np.float64)             # math convention, m[row, col]
    y_lo, y_hi = 2.0 * (y0 - 2.0) / height - 1.0, 2.0 * (y1 + 2.0) / height - 1.0
    planes = [m[3] + m[0], m[3] - m[0], m[1] - y_lo * m[3], y_hi * m[3] - m[1]]
    planes = [p.astype(np.float32).astype(np.float64) for p in planes]
    return [(p, float(np.float32(np.linalg.norm(p[:3]) * 1.0001))) for p in planes]


def chunk_is_rejected(centre, radius, inst, planes):
    ts, rot = inst["translation_and_scale"].astype(np.float64), inst["rotation"].astype(np.float64)
    wc = ts[:3] + rotate(rot, centre) * ts[3]
    r = abs(radius * ts[3]) * 1.01 + 1e-3
    for p, norm in planes:
        d = p[0] * wc[0] + p[1] * wc[1] + p[2] * wc[2] + p[3]
        slack = r * norm + 1e-4 * (abs(p[0] * wc[0]) + abs(p[1] * wc[1]) + abs(p[2] * wc[2]) + abs(p[3]))
        if d < -slack:
            return True
    return False


def random_scene(rng, w, h):
    cam = scenes.Camera(w, h, tuple(rng.uniform([-3, 0.5, -3], [3, 6, 3])), float(rng.uniform(-180, 180)), float(rng.uniform(-50, 30)))
    meshes = scenes.MeshSet()
    kinds = [lambda: scenes.uv_sphere(32, 16), scenes.box_mesh, lambda: scenes.torus_knot(n_u=96, n_v=16), lambda: scenes.quad_mesh(1.0),
             lambda: scenes.uv_sphere(64, 32)]
    inst = []
    for _ in range(int(rng.integers(20, 60))):
        prim = meshes.add(kinds[int(rng.integers(0, len(kinds)))](), int(rng.choice([0, 0, 2])))
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        # near the camera, far away, huge (the camera may sit inside), tiny
        scale = float(10.0 ** rng.uniform(-1.5, 1.3))
        inst.append(scenes.make_instance(tuple(cam.position + rng.normal(size=3) * 10.0 ** rng.uniform(-0.5, 1.3)), scale, tuple(q), prim, 0))
    mesh, prims = meshes.arrays()
    return cam, mesh, prims, np.concatenate(inst)


@pytest.mark.parametrize("seed", range(12))
def test_rejected_chunks_contribute_no_pixel_of_the_band(oracle, seed):
    rng = np.random.default_rng(0xC0FFEE + seed)
    w, h = int(rng.integers(96, 320)), int(rng.integers(64, 200))
    cam, mesh, prims, inst = random_scene(rng, w, h)
    n_bands = int(rng.choice([1, 2, 4, 8]))
    band = int(rng.integers(0, n_bands))
    y0, y1 = (band * h) // n_bands, ((band + 1) * h) // n_bands
    planes = band_planes(cam.proj_view, h, y0, y1)
    idx = mesh["indices"].copy()
    keep = np.ones(len(idx) // 3, bool)
    rejected = total = 0
    for i in inst:
        p = prims[int(i["primitive_id"])]
        first, count = int(p["first_index"]), int(p["index_count"])
        for k, (c, r) in enumerate(chunk_spheres(mesh["positions"], idx[first:first + count])):
            total += 1
            if chunk_is_rejected(c, r, i, planes):
                rejected += 1
                t0 = first // 3 + k * CHUNK
                keep[t0:min(t0 + CHUNK, (first + count) // 3)] = False
    # the filtered mesh: rejected triangles become degenerate (same vertex three times: no area, no pixel), so that every
    # other triangle keeps its id and the tie rule of the rasteriser sees the same order
    filtered = idx.reshape(-1, 3).copy()
    filtered[~keep] = filtered[~keep][:, :1]
    visible = np.arange(len(inst), dtype=np.uint32)      # no instance culling here: the chunk test stands alone
    pc = cam.push_constants()
    full = oracle.visibility(mesh, inst, prims, visible, pc, y0, y1)
    part = oracle.visibility(dict(mesh, indices=filtered.reshape(-1)), inst, prims, visible, pc, y0, y1)
    for layer, (a, b) in enumerate(zip(full, part)):
        for k in ("depth", "normal", "uv", "material_id") + (("scale",) if layer == 1 else ()):
            assert a[k][y0:y1].tobytes() == b[k][y0:y1].tobytes(), (seed, layer, k, rejected, total)
    assert total > 50
    if n_bands > 1:
        assert rejected > 0, "a strict band always has chunks above or below it in these scenes"
    covered = (full[0]["depth"][y0:y1] > 0).mean() + (full[1]["depth"][y0:y1] > 0).mean()
    assert covered > 0.02, "the band shows something"


def instance_on_band(sphere, inst, proj_view, height, y0, y1):
    """band_filter_kernel's instance_on_band: conservative rows of the instance's bounding sphere, clip(C + d) = clip(C) + M d."""
    m = proj_view.astype(np.float64)
    ts, rot = inst["translation_and_scale"].astype(np.float64), inst["rotation"].astype(np.float64)
    c = ts[:3] + rotate(rot, sphere[:3].astype(np.float64)) * ts[3]
    cc = m @ np.append(c, 1.0)
    r = abs(float(sphere[3]) * ts[3]) * 1.01 + 1e-3
    ny, nw = np.linalg.norm(m[1, :3]) * 1.0001, np.linalg.norm(m[3, :3]) * 1.0001
    w_lo, w_hi = cc[3] - r * nw, cc[3] + r * nw
    if not w_lo > 0.0:
        return True
    y_lo, y_hi = cc[1] - r * ny, cc[1] + r * ny
    n_lo, n_hi = min(y_lo / w_lo, y_lo / w_hi), max(y_hi / w_lo, y_hi / w_hi)
    return (n_hi + 1.0) * 0.5 * height + 2.0 >= y0 and (n_lo + 1.0) * 0.5 * height - 2.0 <= y1


@pytest.mark.parametrize("seed", range(12))
def test_band_filter_keeps_every_instance_that_reaches_the_band(oracle, seed):
    """The band's own work list (band_filter_kernel) drops an instance only when its bounding sphere cannot reach the band's
    rows: rendering the band from the kept instances alone gives the same bytes.  The spheres are the primitives' true bounds,
    as the reference's loader computes them (src/model_loading.rs:148-155) and as its own frustum cull relies on."""
    rng = np.random.default_rng(0xBA2D + seed)
    w, h = int(rng.integers(96, 320)), int(rng.integers(64, 200))
    cam, mesh, prims, inst = random_scene(rng, w, h)
    n_bands = int(rng.choice([2, 4, 8]))
    band = int(rng.integers(0, n_bands))
    y0, y1 = (band * h) // n_bands, ((band + 1) * h) // n_bands
    # true bounding spheres (centre of the box, farthest vertex)
    prims = prims.copy()
    for p in prims:
        v = mesh["positions"][mesh["indices"][int(p["first_index"]):int(p["first_index"]) + int(p["index_count"])]].astype(np.float64)
        c = 0.5 * (v.min(0) + v.max(0))
        p["packed_bounding_sphere"] = (*c, np.sqrt(((v - c) ** 2).sum(1).max()) * (1 + 1e-6))
    every = np.arange(len(inst), dtype=np.uint32)
    kept = np.array([i for i in every if instance_on_band(prims[int(inst[i]["primitive_id"])]["packed_bounding_sphere"], inst[i],
                                                            cam.proj_view, h, y0, y1)], dtype=np.uint32)
    pc = cam.push_constants()
    full = oracle.visibility(mesh, inst, prims, every, pc, y0, y1)
    part = oracle.visibility(mesh, inst, prims, kept, pc, y0, y1)
    for layer, (a, b) in enumerate(zip(full, part)):
        for k in ("depth", "normal", "uv", "material_id") + (("scale",) if layer == 1 else ()):
            assert a[k][y0:y1].tobytes() == b[k][y0:y1].tobytes(), (seed, layer, k, len(kept), len(every))
    assert 0 < len(kept) < len(every), "a strict band drops some instances and keeps some in these scenes"


def test_coarse_edge_test_never_rejects_an_inside_pixel():
    """The tile pass drops a pixel when an fp32 edge function is below -ebound (raster_tiles_kernel: a = (float)A, b = (float)B,
    c = (float)(A X0 + B Y0 + C) at the tile's first pixel centre, value = fmaf(a, x, fmaf(b, y, c)) for x, y in 0..63,
    ebound = (2^-21 (64 (|A| + |B|) + |c|) + 1e-15 (16384 (|A| + |B|) + |C|)) * 1.0001).  The exact rule evaluates
    A px + B py + C in double at absolute pixel centres; wherever that is >= 0 the fp32 test must pass, and the fp32 value
    must stay within ebound of it — over edge coefficients of ten decades, tiles anywhere in a 16 k frame, edges through the tile."""
    rng = np.random.default_rng(0xED6E)
    n = 400_000
    mag = 10.0 ** rng.uniform(-6, 4, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    A, B = mag * np.cos(ang), mag * np.sin(ang)
    A[rng.random(n) < 0.05] = 0.0
    B[rng.random(n) < 0.05] = 0.0
    tx, ty = rng.integers(0, 256, n) * 64.0, rng.integers(0, 256, n) * 64.0
    X0, Y0 = tx + 0.5, ty + 0.5
    # edges that pass near a random point of the tile (the interesting case), and some that do not
    qx, qy = X0 + rng.uniform(-8, 72, n), Y0 + rng.uniform(-8, 72, n)
    C = -(A * qx + B * qy) + np.where(rng.random(n) < 0.2, mag * 10.0 ** rng.uniform(-3, 3, n) * rng.choice([-1, 1], n), 0.0)
    x, y = rng.integers(0, 64, n).astype(np.float64), rng.integers(0, 64, n).astype(np.float64)
    cl = A * X0 + B * Y0 + C
    a, b, c = A.astype(np.float32), B.astype(np.float32), cl.astype(np.float32)
    f32 = lambda v: v.astype(np.float32).astype(np.float64)   # noqa: E731 - round to binary32
    v = f32(b.astype(np.float64) * y + c.astype(np.float64))                 # fmaf(b, y, c): one rounding
    value = f32(a.astype(np.float64) * x + v)                                # fmaf(a, x, v)
    m = 64.0 * (np.abs(A) + np.abs(B)) + np.abs(cl)
    m_abs = 16384.0 * (np.abs(A) + np.abs(B)) + np.abs(C)
    ebound = ((m * 4.76837158203125e-7 + m_abs * 1e-15).astype(np.float32) * np.float32(1.0001)).astype(np.float64)
    exact = A * (tx + x + 0.5) + B * (ty + y + 0.5) + C                      # the double rule at the absolute pixel centre
    assert (np.abs(value - exact) <= ebound).all(), float((np.abs(value - exact) / np.maximum(ebound, 1e-300)).max())
    inside = exact >= 0.0
    assert 0.2 < inside.mean() < 0.8
    assert not (value[inside] < -ebound[inside]).any()
    # and the bound is not vacuous: it is a small fraction of the edge function's range over the tile
    assert np.median(ebound / np.maximum(64.0 * (np.abs(A) + np.abs(B)), 1e-300)) < 1e-5

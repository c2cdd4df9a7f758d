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


def test_conservative_depth_plane_bounds_the_exact_depth():
    """The tile pass tests an fp32 depth plane d0 + gx x + gy y (+ margin) against the tile's current depth before it evaluates a
    pixel exactly (raster_tiles_kernel).  For every pixel centre inside the triangle the exact rule's depth (eval_pixel: double
    edge functions, fp32 barycentrics, fp32 z / w) must not exceed plane + margin — otherwise a visible fragment could be dropped —
    and must not fall below plane - margin either (the bound is two-sided).  Random front-facing triangles from sub-pixel to
    tile-sized, vertex w over four decades (strong perspective), anywhere in a 16 k frame."""
    rng = np.random.default_rng(0xDE97)
    f32 = lambda v: np.asarray(v, np.float64).astype(np.float32).astype(np.float64)   # noqa: E731 - round to binary32
    checked = worst = 0.0
    for _ in range(4000):
        tile_x0, tile_y0 = int(rng.integers(0, 256)) * 64, int(rng.integers(0, 256)) * 64
        size = 10.0 ** rng.uniform(-0.5, 2.0)
        centre = np.array([tile_x0, tile_y0]) + rng.uniform(0, 64, 2)
        pix = centre + rng.normal(size=(3, 2)) * size
        W = f32(10.0 ** rng.uniform(-1.3, 2.7, 3) if rng.random() < 0.5 else 10.0 ** rng.uniform(-1.3, 2.7) * rng.uniform(0.8, 1.25, 3))
        depth = rng.uniform(0.0005, 0.999, 3)
        sx, sy, Z = f32(pix[:, 0] * W), f32(pix[:, 1] * W), f32(depth * W)
        det = sx[0] * (sy[1] * W[2] - W[1] * sy[2]) + sy[0] * (W[1] * sx[2] - sx[1] * W[2]) + W[0] * (sx[1] * sy[2] - sy[1] * sx[2])
        if det == 0.0:
            continue
        order = [0, 2, 1] if det < 0.0 else [0, 1, 2]       # setup_front keeps det < 0 and reorders (0, 2, 1)
        rx, ry, Z, W = sx[order], sy[order], Z[order], W[order]
        A = np.array([ry[(i + 1) % 3] * W[(i + 2) % 3] - W[(i + 1) % 3] * ry[(i + 2) % 3] for i in range(3)])
        B = np.array([W[(i + 1) % 3] * rx[(i + 2) % 3] - rx[(i + 1) % 3] * W[(i + 2) % 3] for i in range(3)])
        C = np.array([rx[(i + 1) % 3] * ry[(i + 2) % 3] - ry[(i + 1) % 3] * rx[(i + 2) % 3] for i in range(3)])
        cl = A * (tile_x0 + 0.5) + B * (tile_y0 + 0.5) + C
        dets, absdet = float((cl * W).sum()), float(np.abs(cl * W).sum())
        if not (W.min() > 0.0 and dets > 0.0 and absdet < dets * 1048576.0):
            continue                                          # the kernel switches the plane off (margin = inf) here
        r = 1.0 / dets
        d0, gx, gy = f32((cl * Z).sum() * r), f32((A * Z).sum() * r), f32((B * Z).sum() * r)
        mg = float(np.float32(np.float32(abs(d0) + 64.0 * (abs(gx) + abs(gy))) * np.float32(1.9073486e-6)
                              + np.float32(np.abs(Z).max() / W.min()) * np.float32(9.536743e-7)))
        y, x = np.mgrid[0:64, 0:64]
        qx, qy = tile_x0 + x + 0.5, tile_y0 + y + 0.5
        E = A[:, None, None] * qx + B[:, None, None] * qy + C[:, None, None]
        S = E.sum(0)
        inside = (E >= 0.0).all(0) & (S > 0.0)
        if not inside.any():
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            l = f32(E * (1.0 / S))
            zq = f32(f32(f32(l[0] * Z[0]) + f32(l[1] * Z[1])) + f32(l[2] * Z[2]))
            wq = f32(f32(f32(l[0] * W[0]) + f32(l[1] * W[1])) + f32(l[2] * W[2]))
            d = f32(zq / wq)
        plane = f32(gx * x + f32(gy * y + d0))               # fmaf(gx, x, fmaf(gy, y, d0))
        ok = inside & (wq > 0.0)
        err = np.abs(plane - d)[ok]
        assert (err <= mg).all(), (float(err.max()), mg, size, W.tolist())
        checked += ok.sum()
        worst = max(worst, float((err / mg).max()))
    assert checked > 200_000
    assert worst > 1e-3, "the margin is not orders of magnitude larger than the errors it covers"

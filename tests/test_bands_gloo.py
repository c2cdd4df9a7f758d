"""world_size-2 `gloo` test (CPU) of the multi-GPU band path's host logic: the unique id and IPC handles reach every
rank, the bands partition the frame, and band-wise shading stitched together equals the full frame bit for bit
(the per-pixel path has no cross-band dependency except the opaque frame read by the refraction fetch)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from transmission_renderer_b200 import host, parallel, scenes

W, H = 96, 54


class FakeRenderer:
    """Records what the band plumbing hands to the C ABI (tr_comm_init / tr_peer_attach); no GPU involved."""
    height = H

    def __init__(self, rank):
        self.rank = rank
        self.calls = []

    @staticmethod
    def comm_unique_id():
        return bytes(range(128))

    def comm_init(self, uid, rank, world):
        self.calls.append(("comm_init", uid, rank, world))

    def peer_export(self):
        return bytes([self.rank]) * 64

    def peer_attach(self, rank, world, handles):
        self.calls.append(("peer_attach", rank, world, b"".join(handles)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as oracle
        from pipeline import oracle_cluster_lights, oracle_scene
        r = FakeRenderer(rank)
        y0, y1 = parallel.init_bands(r, rank, world, exchange="peer")
        assert r.calls[0] == ("comm_init", bytes(range(128)), rank, world)
        assert r.calls[1] == ("peer_attach", rank, world, b"".join(bytes([k]) * 64 for k in range(world)))
        cam = scenes.Camera(W, H, (0.0, 2.0, 5.0), 0.0, -5.0)
        centres = [(-1.0, 1.5, 0.0), (1.2, 2.0, -1.0), (0.0, 0.6, 1.0)]
        g0 = scenes.raycast_spheres(cam, centres, [1.0, 0.8, 0.5], [0, 1, 2])
        mats = scenes.hashed_materials(3, 7)
        lights = scenes.config2_lights()
        uniforms = host.make_uniforms(W, H)
        _, cc, ci = oracle_cluster_lights(oracle, cam, uniforms, lights)
        sc = oracle_scene(cam.push_constants(), uniforms, mats, lights, cc, ci)
        _, band16 = oracle.shade_opaque_frame(g0, sc, y0, y1)
        full = parallel.gather_bands(band16[y0:y1], H, rank, world)
        _, ref16 = oracle.shade_opaque_frame(g0, sc)
        assert full.shape == ref16.shape and full.tobytes() == ref16.tobytes()
        out.put((rank, y0, y1))
    finally:
        dist.destroy_process_group()


def test_band_rows_partition_the_frame():
    for h in (1080, 2160, 4320, 187, 7):
        for n in (1, 2, 3, 4, 8):
            rows = [parallel.band_rows(h, r, n) for r in range(n)]
            assert rows[0][0] == 0 and rows[-1][1] == h
            assert all(rows[i][1] == rows[i + 1][0] for i in range(n - 1))


@pytest.mark.timeout(180)
def test_two_rank_bands_equal_full_frame():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert got == [(0, 0, H // 2), (1, H // 2, H)]


def _group_worker(rank, world, port, out):
    """bench.py's views x bands layout: 2 view groups of 2 band ranks, one gloo subgroup per view group."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bands = 2
        group_id, band_rank = rank // bands, rank % bands
        mine = None
        for g in range(world // bands):
            sub = dist.new_group(ranks=list(range(g * bands, (g + 1) * bands)), backend="gloo")
            if g == group_id:
                mine = sub

        class GroupRenderer(FakeRenderer):
            @staticmethod
            def comm_unique_id():
                return bytes([group_id]) * 128     # each view group has its own communicator

        r = GroupRenderer(band_rank)
        y0, y1 = parallel.init_bands(r, band_rank, bands, group=mine, exchange="peer")
        assert r.calls[0] == ("comm_init", bytes([group_id]) * 128, band_rank, bands)
        assert r.calls[1] == ("peer_attach", band_rank, bands, b"".join(bytes([k]) * 64 for k in range(bands)))
        full = parallel.gather_bands(np.full((y1 - y0, 4, 1), rank, np.uint8), H, band_rank, bands, group=mine)
        assert full.shape == (H, 4, 1) and set(np.unique(full)) == {group_id * bands, group_id * bands + 1}
        out.put((rank, group_id, y0, y1))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_view_groups_of_band_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_group_worker, args=(r, 4, port, out)) for r in range(4)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(4))
    assert got == [(0, 0, 0, H // 2), (1, 0, H // 2, H), (2, 1, 0, H // 2), (3, 1, H // 2, H)]


# ---- cost-balanced bands (tr_set_bands / parallel.balance_bands) ---------------------------------------------------------
def test_balanced_bounds_properties():
    rng = np.random.default_rng(0)
    for n in (2, 4, 8):
        for h in (1080, 2160, 4320):
            bounds = [parallel.band_rows(h, r, n)[0] for r in range(n)] + [h]
            ms = rng.uniform(0.2, 1.0, n)
            new = parallel.balanced_bounds(bounds, ms)
            assert new[0] == 0 and new[-1] == h and all(b > a for a, b in zip(new, new[1:]))
            assert all(v % 8 == 0 for v in new[1:-1])
            # the slowest band gets fewer rows, the fastest more
            rows_old, rows_new = np.diff(bounds), np.diff(new)
            assert rows_new[np.argmax(ms)] <= rows_old[np.argmax(ms)] and rows_new[np.argmin(ms)] >= rows_old[np.argmin(ms)]
            # equal cost -> nothing moves (up to the alignment)
            same = parallel.balanced_bounds(bounds, np.ones(n))
            assert max(abs(a - b) for a, b in zip(same, bounds)) <= 4


def test_balanced_bounds_converge_on_a_known_cost_profile():
    """Cost per row 1 + 3 y / H: repeated balancing against the exact band costs ends with equal-cost bands."""
    h, n = 2160, 8
    cost = 1.0 + 3.0 * np.arange(h) / h
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [parallel.band_rows(h, r, n)[0] for r in range(n)] + [h]
    for _ in range(6):
        bounds = parallel.balanced_bounds(bounds, [cum[b] - cum[a] for a, b in zip(bounds, bounds[1:])])
    ms = np.array([cum[b] - cum[a] for a, b in zip(bounds, bounds[1:])])
    assert ms.max() / ms.mean() < 1.03


class TimedFakeRenderer(FakeRenderer):
    """A band's cost grows with its rows and with how far down the frame it lies."""

    def __init__(self, rank, world):
        super().__init__(rank)
        self.y0, self.y1 = parallel.band_rows(H * 40, rank, world)
        self.height = H * 40
        self.frames = 0

    def enable_timing(self, on):
        self.frames = 0

    def sync(self):
        pass

    def render(self):
        self.frames += 1

    def pass_totals(self):
        ys = np.arange(self.y0, self.y1)
        ms = float((1.0 + 3.0 * ys / self.height).sum()) * 1e-3
        return {"visibility_ms": ms * self.frames, "shade_opaque_ms": 0.0, "shade_transmission_ms": 0.0, "tonemap_ms": 0.0}, self.frames

    def set_bands(self, bounds):
        self.calls.append(("set_bands", tuple(bounds)))
        self.y0, self.y1 = bounds[self.rank], bounds[self.rank + 1]


def _balance_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r = TimedFakeRenderer(rank, world)
        bounds = parallel.balance_bands(r, lambda: [r.render() for _ in range(3)], rank, world, iterations=5)
        r.frames = 1
        out.put((rank, tuple(bounds), r.pass_totals()[0]["visibility_ms"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_balance_bands_agree_and_equalise():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_balance_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert got[0][1] == got[1][1]                      # every rank computed the same boundaries
    assert got[0][1][1] > (H * 40) // 2                # the cheaper top band grew
    assert abs(got[0][2] - got[1][2]) / max(got[0][2], got[1][2]) < 0.05

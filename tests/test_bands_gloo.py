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

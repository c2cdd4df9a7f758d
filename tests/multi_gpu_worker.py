"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run, one rank per GPU.  Renders the same small
frame (a) on every rank alone, full frame, and (b) band-sharded across the ranks with the NCCL all-gather and with
the fused peer-store exchange, and checks (b) == (a) bit for bit on every rank's band."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transmission_renderer_b200 import Renderer, host, parallel, scenes  # noqa: E402


def render(s, lut, w, h, rank, world, group, exchange, frames=3, ray_tracing=False, bounds=None):
    cam = s["camera"]
    with Renderer(w, h, device=int(os.environ.get("LOCAL_RANK", "0"))) as r:
        r.set_uniforms(s["uniforms"]); r.set_materials(s["materials"]); r.set_lights(s["lights"]); r.set_ggx_lut(lut)
        r.set_instances(s["instances"]); r.set_primitives(s["primitives"])
        m = s["mesh"]
        r.set_mesh(m["positions"], m["normals"], m["uvs"], m["indices"])
        r.build_clusters(cam.write_cluster_data())
        y0, y1 = parallel.init_bands(r, rank, world, group=group, exchange=exchange) if world > 1 else (0, h)
        if bounds is not None:   # caller-balanced (ragged) bands, tr_set_bands
            r.set_bands(bounds)
            y0, y1 = bounds[rank], bounds[rank + 1]
        handle = r.build_acceleration_structures() if ray_tracing else 0   # every rank builds the same structures
        fp = cam.frame_params(host.default_tonemap_params(), acceleration_structure_address=handle)
        for _ in range(frames):   # several frames back to back: the exchange buffers are reused
            r.frame(fp)
        r.sync()
        hdr = r.read_hdr()[y0:y1].copy()
        srgb = r.read_srgb8()[y0:y1].copy()
        mip0 = r.read_pyramid_level(0).copy()
        mip3 = r.read_pyramid_level(3).copy()
    return (y0, y1), hdr, srgb, mip0, mip3


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    group = dist.new_group(backend="gloo")
    w, h = 640, 362   # 362 rows: unequal bands for 4 ranks
    z = np.load(os.path.join(ROOT, "tests", "golden", "ggx_lut_rg.npz"))
    lut = np.zeros(z["rg"].shape[:2] + (4,), np.uint8)
    lut[..., :2] = z["rg"]
    lut[..., 3] = 255
    s = scenes.instanced_scene(w, h, n_instances=3000, n_lights=32)
    _, hdr1, srgb1, mip0_1, mip3_1 = render(s, lut, w, h, 0, 1, None, "nccl")
    for exchange in ("nccl", "peer"):
        (y0, y1), hdr, srgb, mip0, mip3 = render(s, lut, w, h, rank, world, group, exchange)
        assert hdr.tobytes() == hdr1[y0:y1].tobytes(), f"{exchange}: HDR band differs from the single-GPU frame"
        assert srgb.tobytes() == srgb1[y0:y1].tobytes(), f"{exchange}: sRGB8 band differs"
        assert mip0.tobytes() == mip0_1.tobytes(), f"{exchange}: exchanged opaque frame differs"
        assert mip3.tobytes() == mip3_1.tobytes(), f"{exchange}: mip 3 differs"
        dist.barrier(group=group)
    # caller-balanced bands (tr_set_bands): ragged boundaries, both exchanges
    cuts = sorted({int(round(h * f / 8.0)) * 8 for f in (0.17, 0.49, 0.8)})[: world - 1]
    bounds = [0] + cuts + [h]
    assert len(bounds) == world + 1
    for exchange in ("nccl", "peer"):
        (y0, y1), hdr, srgb, mip0, mip3 = render(s, lut, w, h, rank, world, group, exchange, bounds=bounds)
        assert (y0, y1) == (bounds[rank], bounds[rank + 1])
        assert hdr.tobytes() == hdr1[y0:y1].tobytes(), f"{exchange}, balanced bands {bounds}: HDR band differs from the single-GPU frame"
        assert srgb.tobytes() == srgb1[y0:y1].tobytes() and mip0.tobytes() == mip0_1.tobytes() and mip3.tobytes() == mip3_1.tobytes()
        dist.barrier(group=group)
    # ray-queried shadows: every rank traces the rays of its own rows against the whole scene
    _, hdr1, srgb1, mip0_1, _ = render(s, lut, w, h, 0, 1, None, "nccl", frames=1, ray_tracing=True)
    (y0, y1), hdr, srgb, mip0, _ = render(s, lut, w, h, rank, world, group, "peer", frames=2, ray_tracing=True)
    assert hdr.tobytes() == hdr1[y0:y1].tobytes(), "ray queries: HDR band differs from the single-GPU frame"
    assert srgb.tobytes() == srgb1[y0:y1].tobytes() and mip0.tobytes() == mip0_1.tobytes(), "ray queries: band / exchanged frame differs"
    dist.barrier(group=group)
    if rank == 0:
        print(f"multi-GPU OK: {world} ranks, nccl + peer exchange, equal and balanced bands == single-GPU frame bitwise")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

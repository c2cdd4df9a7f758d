"""Host-side plumbing of the multi-GPU band path (one process per GPU, torch.distributed).

The reference is single-GPU (one VkQueue, src/main.rs:243); band sharding is the new work the north star
defines.  The device side lives in csrc/tr_comm.cu; this module only moves the 128-byte NCCL
unique id and the 64-byte CUDA-IPC handles between ranks and stitches bands for callers who want a whole image.
Every function takes the process group to talk over, so the CPU tests can run it on `gloo`.
"""
import numpy as np

from . import host


def band_rows(height, rank, world_size):
    """Rows [y0, y1) of `rank`: floor(r H / N) .. floor((r+1) H / N) — the rule tr_comm_init applies."""
    return host.band_rows(height, rank, world_size)


def init_bands(renderer, rank, world_size, group=None, exchange="nccl"):
    """Create the communicator of `renderer` and set its band.  exchange = "nccl" (all-gather of the opaque
    bands) or "peer" (opaque shading stores its band into every peer over NVLink; needs CUDA IPC)."""
    import torch.distributed as dist
    if exchange not in ("nccl", "peer"):
        raise ValueError(exchange)
    if world_size == 1:
        return band_rows(renderer.height, 0, 1)
    uid = [type(renderer).comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, group=group)
    renderer.comm_init(uid[0], rank, world_size)
    if exchange == "peer":
        handles = [None] * world_size
        dist.all_gather_object(handles, renderer.peer_export(), group=group)
        renderer.peer_attach(rank, world_size, handles)
    return band_rows(renderer.height, rank, world_size)


def gather_bands(band, height, rank, world_size, group=None):
    """band: this rank's rows (y1 - y0, W, C).  Returns the whole (H, W, C) image on every rank."""
    import torch.distributed as dist
    band = np.ascontiguousarray(band)
    if world_size == 1:
        return band
    parts = [None] * world_size
    dist.all_gather_object(parts, band, group=group)
    for r, p in enumerate(parts):
        y0, y1 = band_rows(height, r, world_size)
        if p.shape[0] != y1 - y0:
            raise ValueError(f"rank {r} sent {p.shape[0]} rows for band [{y0},{y1})")
    return np.concatenate(parts, axis=0)

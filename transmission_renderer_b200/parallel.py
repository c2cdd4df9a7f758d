"""Host-side plumbing of the multi-GPU band path (one process per GPU, torch.distributed).

The reference is single-GPU (one VkQueue, src/main.rs:243); band sharding is the new work the north star
defines.  The device side lives in csrc/tr_comm.cu; this module only moves the 128-byte NCCL
unique id and the 64-byte CUDA-IPC handles between ranks and stitches bands for callers who want a whole image.
Every function takes the process group to talk over, so the CPU tests can run it on `gloo`.
"""
import os
import sys

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
    # `rank` counts inside `group` (a view group's bands); broadcast_object_list names its source by GLOBAL rank
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast_object_list(uid, src=src, group=group)
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


def balanced_bounds(bounds, band_ms, align=8, damping=0.7):
    """New band boundaries from the measured time of every band.

    `bounds`: the n + 1 current row boundaries, `band_ms`: what each band cost.  The cost is taken as uniform inside a band,
    which makes the cumulative cost piecewise linear in the row; the new boundaries cut it into n equal parts, moved only
    `damping` of the way (the cost inside a band is not really uniform) and rounded to `align` rows.  Pure function of its
    arguments, so every rank computes the same boundaries from the gathered times."""
    b = np.asarray(bounds, np.float64)
    t = np.maximum(np.asarray(band_ms, np.float64), 1e-9)
    n = len(t)
    cum = np.concatenate([[0.0], np.cumsum(t)])
    target = cum[-1] * np.arange(1, n) / n
    new = np.interp(target, cum, b)                       # rows at which the cumulative cost reaches k / n of the total
    new = b[1:-1] + damping * (new - b[1:-1])
    new = np.round(new / align) * align
    out = np.concatenate([[b[0]], new, [b[-1]]]).astype(np.int64)
    for k in range(1, n):                                 # keep every band at least `align` rows high
        out[k] = min(max(out[k], out[k - 1] + align), int(b[-1]) - (n - k) * align)
    return [int(v) for v in out]


def balance_bands(renderer, run_frames, rank, world_size, group=None, iterations=4):
    """Move the band boundaries until every rank's band costs about the same (tr_set_bands).

    Equal-row bands leave the ranks unequal work where coverage, triangle density and light counts vary down the frame — the
    frame time is the slowest band's.  `run_frames()` renders a few frames with the per-pass timers on; the band-local passes
    (visibility, both shading passes, tonemap) are the band's cost.  Returns the final boundaries."""
    import torch.distributed as dist
    h = renderer.height
    bounds = [band_rows(h, r, world_size)[0] for r in range(world_size)] + [h]
    if world_size == 1:
        return bounds
    log = os.environ.get("TR_BALANCE_LOG")
    for it in range(iterations):
        run_frames()             # untimed: the first frames after new boundaries grow buffers (cudaMalloc inside a pass)
        renderer.sync()
        renderer.enable_timing(True)
        run_frames()
        renderer.sync()
        totals, n = renderer.pass_totals()
        renderer.enable_timing(False)
        mine = sum(totals[k] for k in ("visibility_ms", "shade_opaque_ms", "shade_transmission_ms", "tonemap_ms")) / max(n, 1)
        every = [None] * world_size
        dist.all_gather_object(every, float(mine), group=group)
        if log and rank == 0:
            print(f"balance_bands {it}: rows {bounds} cost ms {[round(v, 3) for v in every]}", file=sys.stderr)
        bounds = balanced_bounds(bounds, every)
        renderer.set_bands(bounds)
    return bounds

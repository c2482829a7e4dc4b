"""Strip sharding helpers: the panorama is cut into N horizontal strips on the 2^nb grid, one per
rank; each rank composes its strip (plus halo) and the finished rows are gathered on rank 0.  The only collective
on the path is that final gather (grouped point-to-point sends of contiguous row blocks: NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import strip_rows


def all_strip_rows(padded_h, final_h, num_bands, world):
    """[(y0, y1)] of every rank - pure arithmetic, identical on all ranks (no communication needed)."""
    return [strip_rows(padded_h, final_h, num_bands, r, world) for r in range(world)]


def gather_strips(tensors, rows, rank, world, dst=0):
    """Gather rows [y0, y1) of every tensor in `tensors` (each indexed [row, ...], same on all ranks) to `dst`.

    `rows[r]` is rank r's (y0, y1).  Full-width row blocks are contiguous, so they are sent in place, without staging.
    Returns after the receives have completed on `dst` (stream-ordered on NCCL)."""
    if world == 1:
        return
    ops = []
    if rank == dst:
        for r in range(world):
            y0, y1 = rows[r]
            if r == dst or y1 <= y0:
                continue
            for t in tensors:
                ops.append(dist.P2POp(dist.irecv, t[y0:y1], r))
    else:
        y0, y1 = rows[rank]
        if y1 > y0:
            for t in tensors:
                ops.append(dist.P2POp(dist.isend, t[y0:y1], dst))
    if ops:  # one grouped launch (ncclGroupStart/End): all strips move concurrently over NVLink
        for q in dist.batch_isend_irecv(ops):
            q.wait()

"""CPU, world_size 2 and 3 over gloo: the N > 1 path's host logic - strip partition + gather of finished rows on rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, padded_h, final_h, nb, width, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from image_stitching_b200 import strips
    rows = strips.all_strip_rows(padded_h, final_h, nb, world)
    # every rank "composes" only its own rows: value = 1 + row index (+ channel), everything else stays 0
    pano = torch.zeros((final_h, width, 3), dtype=torch.uint8)
    mask = torch.zeros((final_h, width), dtype=torch.uint8)
    y0, y1 = rows[rank]
    yy = torch.arange(y0, y1, dtype=torch.int64)
    pano[y0:y1] = ((yy[:, None, None] + torch.arange(3)[None, None, :] + 1) % 251).to(torch.uint8)
    mask[y0:y1] = 255
    strips.gather_strips([pano, mask], rows, rank, world)
    if rank == 0:
        yy = torch.arange(final_h, dtype=torch.int64)
        want = ((yy[:, None, None] + torch.arange(3)[None, None, :] + 1) % 251).to(torch.uint8).expand(final_h, width, 3)
        q.put((bool(torch.equal(pano, want)), bool((mask == 255).all()), rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,padded_h,final_h,nb", [(2, 2912, 2881, 5), (3, 192, 170, 5), (2, 64, 40, 5)])
def test_gather_over_gloo(world, padded_h, final_h, nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, padded_h, final_h, nb, 37, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok_pano, ok_mask, rows = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_pano and ok_mask
    assert rows[0][0] == 0 and rows[-1][1] == final_h and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))

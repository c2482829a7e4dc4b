"""GPU x2 (skipped on a single-GPU box; tests/test_gpu_strips.py runs the same code paths with logical strips on one GPU):
strip sharding across two processes reproduces the ORACLE's panorama bit for bit, with the strips gathered on rank 0 by the
copy engine (local staging + peer copy) and by peer stores from the final kernel."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import image_stitching_b200 as isb
    from conftest import make_case, seam_masks_oracle
    rig, imgs, gains, nb = make_case("cfg3", 16, 3)
    seams = seam_masks_oracle(rig)
    cams = isb.cameras_from_KR(rig.Ks, rig.Rs)
    from oracle import oracle as orc
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams) if rank == 0 else None
    results = []
    for mode in (isb.GATHER_COPY_ENGINE, isb.GATHER_PEER_STORES):
        gm = isb.GATHER_LOCAL if (rank == 0 and mode == isb.GATHER_COPY_ENGINE) else mode
        c = isb.Composer(rig.warp, rig.scale, nb, strip_index=rank, strip_count=world, gather_mode=gm)
        _, _, roi = c.plan(cams, [(rig.W, rig.H)] * rig.n)
        h, w = roi[3], roi[2]
        if rank == 0:
            pano, mask = isb.DevPtr.alloc((h, w, 3)), isb.DevPtr.alloc((h, w))
            handles = [pano.ipc_handle(), mask.ipc_handle()]
        else:
            handles = [None, None]
        dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            pano, mask = isb.DevPtr.open_ipc(handles[0], (h, w, 3)), isb.DevPtr.open_ipc(handles[1], (h, w))
        for _ in range(3):
            r = c.run(imgs, gains, seams, out=pano, out_mask=mask)
        c.sync()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            results.append((bool(np.array_equal(pano.to_numpy(), ref["result8"])), bool(np.array_equal(mask.to_numpy(), ref["mask"])),
                            r["strip_rows"]))
        dist.barrier()
        pano.close()
        mask.close()
    if rank == 0:
        q.put(results)
    dist.destroy_process_group()


def test_two_gpu_strips_with_peer_gather():
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(results) == 2
    for ok8, okm, rows in results:
        assert ok8 and okm and rows[0] == 0

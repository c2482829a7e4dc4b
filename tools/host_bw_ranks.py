"""Aggregate host <-> device bandwidth of a node with every rank copying at once (the ceiling of the N > 1 e2e measurement).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/host_bw_ranks.py

Three placements of the pinned host buffer are compared, every rank moving `MB` up and `MB` down per iteration at once:
  own        each rank's own cudaHostAlloc buffer (first touch by the rank itself)
  shared     one POSIX shared-memory block touched by rank 0, registered by every rank (what bench.py's e2e uses)
  interleave the same block after set_mempolicy(MPOL_INTERLEAVE) over all NUMA nodes in rank 0
Prints the node's NUMA layout, `nvidia-smi topo -m`, and one line per placement.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist

MB = 64


def numa_nodes():
    base = "/sys/devices/system/node"
    try:
        return sorted(int(d[4:]) for d in os.listdir(base) if d.startswith("node") and d[4:].isdigit())
    except OSError:
        return []


def set_interleave(nodes):
    """set_mempolicy(MPOL_INTERLEAVE, mask) through the raw syscall (x86-64: 238); returns errno or 0."""
    libc = ctypes.CDLL(None, use_errno=True)
    mask = 0
    for n in nodes:
        mask |= 1 << n
    m = ctypes.c_ulong(mask)
    rc = libc.syscall(238, 3, ctypes.byref(m), ctypes.c_ulong(64))
    return 0 if rc == 0 else ctypes.get_errno()


def reset_policy():
    libc = ctypes.CDLL(None, use_errno=True)
    libc.syscall(238, 0, None, ctypes.c_ulong(0))


def run(label, host_up, host_down, dev, rank, world, iters=20):
    n = host_up.numel()
    d_up = torch.empty(n, dtype=torch.uint8, device=dev)
    d_down = torch.zeros(n, dtype=torch.uint8, device=dev)
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for timed in (False, True):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s_up.wait_event(e0)
            s_down.wait_event(e0)
            for _ in range(iters if timed else 3):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s_up):
                        d_up.copy_(host_up, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s_down):
                        host_down.copy_(d_down, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_up)
            torch.cuda.current_stream().wait_stream(s_down)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per_dir = n * iters * world / (float(ms.item()) * 1e-3) / 1e9
        res[mode] = per_dir * (2 if mode == "both" else 1)
    if rank == 0:
        print(f"{label:11s} ranks {world}: h2d {res['h2d']:.1f} GB/s, d2h {res['d2h']:.1f} GB/s, both {res['both']:.1f} GB/s (sum of both directions)",
              flush=True)


def shared_block(nbytes, rank, world, interleave):
    from multiprocessing import resource_tracker, shared_memory
    name = [None]
    if rank == 0:
        err = set_interleave(numa_nodes()) if interleave else 0
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        np.ndarray((nbytes,), np.uint8, buffer=shm.buf)[:] = 1  # first touch under the policy
        if interleave:
            reset_policy()
            print(f"set_mempolicy(MPOL_INTERLEAVE) errno {err}", flush=True)
        name[0] = shm.name
    dist.broadcast_object_list(name, src=0)
    if rank != 0:
        shm = shared_memory.SharedMemory(name=name[0])
        try:
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    arr = np.ndarray((nbytes,), np.uint8, buffer=shm.buf)
    rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, nbytes, 0)
    if int(rc) != 0:
        raise RuntimeError(f"cudaHostRegister failed ({rc})")
    return shm, arr


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if rank == 0:
        print("numa nodes:", numa_nodes(), "cpus allowed:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:64], flush=True)
        for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "cat /sys/devices/system/node/node*/cpulist 2>/dev/null"],
                    ["sh", "-c", "grep -E 'MemTotal|MemFree' /sys/devices/system/node/node*/meminfo 2>/dev/null"]):
            try:
                print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout, flush=True)
            except Exception as e:  # noqa: BLE001
                print(cmd, "failed:", e, flush=True)
    n = MB << 20
    up = torch.ones(n, dtype=torch.uint8).pin_memory()
    down = torch.zeros(n, dtype=torch.uint8).pin_memory()
    run("own", up, down, dev, rank, world)
    del up, down
    for label, inter in (("shared", False), ("interleave", True)):
        shm, arr = shared_block(2 * n * world, rank, world, inter)
        up = torch.from_numpy(arr[rank * n:(rank + 1) * n])
        down = torch.from_numpy(arr[(world + rank) * n:(world + rank + 1) * n])
        run(label, up, down, dev, rank, world)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.cudart().cudaHostUnregister(arr.ctypes.data)
        del up, down, arr
        try:
            shm.close()
            if rank == 0:
                shm.unlink()
        except Exception:
            pass
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())

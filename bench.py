#!/usr/bin/env python
"""bench.py - output-panorama MP/s of the fused warp + multi-band blend on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

One "step" = one pass of the compositing hot path (image_stitching.cpp:1086-1229: warp, mask, gain, ->16S,
seam mask, MultiBandBlender prepare/feed/blend, saturate to 8U) over a synthetic rig of SURVEY.md 8(d); the default
is cfg2: 8 x 4000x3000 images, spherical warp, 5 bands, 20912x2881 panorama.

  value  : device-resident inputs -> device-resident panorama (on rank 0 when N > 1), CUDA-event timed, max over ranks.
  e2e    : the same step through the C ABI with pinned HOST buffers (H2D of the sources and D2H of the panorama inside
           the timed region).  N > 1: sources and panorama live in pinned host memory shared by the ranks of the node
           (one copy); every rank uploads the source row bands its strip reads and downloads its strip into the
           shared panorama over its own PCIe link.
  N > 1  : the panorama is cut into N horizontal strips (2^nb-aligned; 4 + 3 halo cells recomputed); each rank composes
           its strip; the strips are gathered into rank 0's double-buffered panorama over NVLink - by the copy engine
           (default, overlaps the next step), by peer stores from the final kernel (--gather p2p) or by NCCL (--gather
           nccl).  scaling = "strong".  Every line carries `parity` (gathered panorama vs the unsharded result computed on
           rank 0, and vs cv2) and a `cfg3` sub-record (the 36 x 24 MP rig the north star names for strip scaling).
  output_side (N = 1): what follows blend() in the reference - imwrite("result.jpg", result) (:1228) and cropper.cpp's crop():
           the JPEG file of the full-size panorama from isb_jpeg_encode, compared byte for byte with cv2.imencode and timed
           against it, and the crop rectangle of the composited mask.
  e2e_to_jpeg (N = 1): decoded frames in pinned host memory -> result.jpg bytes in pinned host memory (compose + encode on
           the device, `--e2e-depth` host threads with one composer / stream each); the file is compared with cv2.imencode's.
  --impl reference : the reference's own CPU implementation of the path (OpenCV's cv::detail classes, driven through cv2
           in the reference's call order) on the host cores, all images of the rig per step; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "output_panorama_MP_per_s_warp_plus_multiband_blend"
UNIT = "MP/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------------------------------
def make_inputs(workload, div=1, images=True):
    from concurrent.futures import ThreadPoolExecutor

    from image_stitching_b200 import synth
    rig = synth.make_rig(workload, scale_div=div)
    imgs = None
    if images:
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:  # numpy releases the GIL in the big ops
            imgs = list(ex.map(lambda i: synth.make_image(i, rig.W, rig.H), range(rig.n)))
    return rig, imgs, synth.make_gains(rig.n)


def torch_images(rig, dev, seed0=1000):
    """Device-generated stand-ins of synth.make_image (same construction: x32 bilinear field + noise in [-10, 10]) for
    workloads whose numpy generation would take minutes (cfg3: 36 x 24 MP).  Only used where the checker is the CUDA
    path itself on the same inputs (strips vs unsharded); oracle / cv2 parity runs use the numpy images."""
    import torch
    out = []
    for i in range(rig.n):
        g = torch.Generator(device=dev)
        g.manual_seed(seed0 + i)
        low = torch.randint(0, 256, (1, 3, rig.H // 32 + 2, rig.W // 32 + 2), generator=g, device=dev, dtype=torch.int32).float()
        up = torch.nn.functional.interpolate(low, scale_factor=32, mode="bilinear", align_corners=False)[0, :, : rig.H, : rig.W]
        up = up.permute(1, 2, 0)
        up = up + torch.randint(-10, 11, up.shape, generator=g, device=dev, dtype=torch.int32).float()
        out.append(up.round_().clamp_(0, 255).to(torch.uint8).contiguous())
        del low, up
    return out


def workload_name(rig, name, div):
    return (f"{name}{'' if div == 1 else '/div' + str(div)}: {rig.n}x{rig.W}x{rig.H} 8UC3, {rig.warp} warp, "
            f"{rig.nb}-band multi-band blend")


def config_of(rig, name, div, roi):
    """The `config` both arms print - identical dicts, so the driver can tell that they ran the same thing."""
    return {"workload": workload_name(rig, name, div), "panorama": [int(roi[2]), int(roi[3])],
            "output_MP": roi[2] * roi[3] / 1e6, "images_per_step": rig.n}


def seam_masks_gpu(rig):
    """Seam-scale auxiliary warp (image_stitching.cpp:973-989) through our warper (bit-exact vs cv2, see tests)."""
    import image_stitching_b200 as isb
    from image_stitching_b200 import synth
    src = synth.seam_source_mask(rig.W, rig.H)
    out = []
    for K, R in zip(rig.Ks, rig.Rs):
        Ks, ss = synth.seam_camera(K, rig.scale)
        out.append(isb.RotationWarper(rig.warp, ss).warp(src, Ks, R, isb.INTER_NEAREST, isb.BORDER_CONSTANT)[1])
    return out


def cpu_reference_step(rig, imgs, gains, seams, timings=None):
    """One pass of the reference CPU path over ALL images of the rig; returns (seconds, output MP, result)."""
    from oracle import cv_reference as cvr
    t0 = time.perf_counter()
    ref = cvr.compose_cv(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, rig.nb, gains, seams, timings=timings)
    dt = time.perf_counter() - t0
    return dt, ref["dst_roi"][2] * ref["dst_roi"][3] / 1e6, ref


# ---------------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    import cv2
    from oracle import cv_reference as cvr
    cv2.ocl.setUseOpenCL(False)
    # IPP OFF in both CPU legs (this arm and the `cpu_baseline` pass of our arm): the reference's vcpkg build has no `ipp`
    # feature (vcpkg.json:6-9) and it is the setting in which the wheel is bit-exact with the restatement
    cv2.ipp.setUseIPP(False)
    rig, imgs, gains = make_inputs(args.workload, args.div)
    seams = cvr.seam_masks_cv(rig.warp, rig.scale, rig.Ks, rig.Rs, rig.W, rig.H)
    steps = args.steps if args.steps is not None else 10
    for _ in range(args.warmup):
        cpu_reference_step(rig, imgs, gains, seams)
    tot, mp, roi = 0.0, 0.0, None
    for _ in range(steps):
        dt, m, ref = cpu_reference_step(rig, imgs, gains, seams)
        tot += dt
        mp += m
        roi = ref["dst_roi"]
    val = mp / tot
    # one more pass with IPP as the wheel ships it, reported beside the headline (only the f32 gain resize changes)
    cv2.ipp.setUseIPP(True)
    dt_ipp, m_ipp, _ = cpu_reference_step(rig, imgs, gains, seams)
    cores = cv2.getNumThreads()
    sample = (f"all {rig.n} of {rig.n} images of {args.workload} ({rig.W}x{rig.H}, {rig.warp}, {rig.nb} bands) per step; "
              f"OpenCV {cv2.__version__} (cv2 wheel, IPP off like the reference's vcpkg build) cv::detail warper/compensator/"
              f"MultiBandBlender in the reference's call order; the reference's main() needs OpenCV dev files + libexif "
              f"and cannot be built here")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8/s16/f32", "data": "synthetic",
            "config": config_of(rig, args.workload, args.div, roi),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(cores), "kind": "reference", "sample": sample,
                             "host_cpus": os.cpu_count(), "ipp": False, "value_with_ipp_on": m_ipp / dt_ipp},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class SharedPinned:
    """Pinned host memory shared by the ranks of one node (POSIX shared memory + cudaHostRegister): the host-side home of
    the sources and of the panorama in the N > 1 end-to-end measurement."""

    def __init__(self, nbytes, rank, world, tag):
        import torch
        import torch.distributed as dist
        from multiprocessing import resource_tracker, shared_memory
        self.nbytes = int(nbytes)
        name = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
            name[0] = self.shm.name
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:  # the creator unlinks; attached processes must not (python < 3.13 registers them with the tracker)
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.rank = rank
        self.arr = np.ndarray((self.nbytes,), np.uint8, buffer=self.shm.buf)
        self.registered = False
        rc = torch.cuda.cudart().cudaHostRegister(self.arr.ctypes.data, self.nbytes, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed ({rc})")
        self.registered = True

    def view(self, offset, shape):
        n = int(np.prod(shape))
        return self.arr[offset:offset + n].reshape(shape)

    def close(self):
        import torch
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.arr.ctypes.data)
            self.registered = False
        self.arr = None
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


def measure(name, div, args, rank, world, dev, stream, full):
    """Times one workload.  full: with e2e, clocks, CPU baseline (the headline workload); else device value + parity only."""
    import torch
    import torch.distributed as dist

    import image_stitching_b200 as isb
    from image_stitching_b200 import strips

    local = dev.index
    rig, imgs, gains = make_inputs(name, div, images=full and not args.device_images)
    seams = seam_masks_gpu(rig)
    cams = isb.cameras_from_KR(rig.Ks, rig.Rs)
    sizes = [(rig.W, rig.H)] * rig.n
    mode = args.gather if world > 1 else "none"
    gm = {"none": isb.GATHER_PEER_STORES, "p2p": isb.GATHER_PEER_STORES, "nccl": isb.GATHER_LOCAL,
          "copy": isb.GATHER_LOCAL if rank == 0 else isb.GATHER_COPY_ENGINE}[mode]
    depth = max(1, args.inflight)

    def make_composer(d):
        c = isb.Composer(rig.warp, rig.scale, rig.nb, strip_index=rank, strip_count=world, cache_plan=True, gather_mode=gm,
                         pipeline_depth=d)
        t0 = time.perf_counter()
        geo = c.plan(cams, sizes)
        torch.cuda.synchronize()
        return c, geo, time.perf_counter() - t0

    comp, (corners, rsizes, roi), plan_s = make_composer(depth)
    pw, ph = roi[2], roi[3]
    out_mp = pw * ph / 1e6

    d_imgs = [torch.from_numpy(im).to(dev) for im in imgs] if imgs is not None else torch_images(rig, dev)
    d_gains = [torch.from_numpy(g).to(dev) for g in gains]
    d_seams = [torch.from_numpy(s).to(dev) for s in seams]
    # every rank's rows (the planner balances the cuts by work; identical on all ranks, gathered here for the NCCL path)
    all_rows = [comp.planned_rows()]
    if world > 1:
        all_rows = [None] * world
        dist.all_gather_object(all_rows, comp.planned_rows())

    # ---- the panorama(s) on rank 0 ----------------------------------------------------------------------------------
    peer = mode in ("p2p", "copy")
    # runs in flight (pipeline depth; copy-engine gather: step k's strips still travel while step k + 1 is composed) write
    # to different panoramas
    nbuf = max(depth, 2 if mode == "copy" else 1)
    # rows padded to 128 B so that staged 16-byte stores / full-sector copies apply over NVLink
    p8, pm = (pw * 3, pw) if not peer else ((pw * 3 + 127) // 128 * 128, (pw + 127) // 128 * 128)
    panos = []
    if peer:
        ok = 1
        handles = [None] * (2 * nbuf)
        try:
            if rank == 0:
                for _ in range(nbuf):
                    panos.append((isb.DevPtr.alloc((ph, p8)), isb.DevPtr.alloc((ph, pm))))
                handles = [h for a, b in panos for h in (a.ipc_handle(), b.ipc_handle())]
        except Exception as e:  # noqa: BLE001
            ok = 0
            sys.stderr.write(f"rank {rank}: peer-memory export failed ({e}); falling back to NCCL gather\n")
        dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            try:
                for k in range(nbuf):
                    panos.append((isb.DevPtr.open_ipc(handles[2 * k], (ph, p8)), isb.DevPtr.open_ipc(handles[2 * k + 1], (ph, pm))))
            except Exception as e:  # noqa: BLE001
                ok = 0
                sys.stderr.write(f"rank {rank}: peer-memory import failed ({e}); falling back to NCCL gather\n")
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not bool(flag.item()):  # every rank takes the same path
            peer, mode, panos, gm = False, "nccl", [], isb.GATHER_LOCAL
            comp, _, _ = make_composer(depth)
    if mode == "nccl":
        depth, nbuf = 1, 1  # the NCCL gather is ordered on the caller's stream: one run at a time
        comp, _, _ = make_composer(1)
    if not peer:
        p8, pm = pw * 3, pw
        panos = [(torch.zeros((ph, pw, 3), dtype=torch.uint8, device=dev), torch.zeros((ph, pw), dtype=torch.uint8, device=dev))
                 for _ in range(nbuf)]
    counter = [0]
    active = [comp]

    def step_device():
        d_out, d_mask = panos[counter[0] % nbuf]
        counter[0] += 1
        active[0].run(d_imgs, d_gains, d_seams, out=d_out, out_mask=d_mask, out_pitch=p8, mask_pitch=pm)
        if mode == "nccl":  # strips -> rank 0 over NCCL (full-width rows are contiguous slices of the panorama tensors)
            strips.gather_strips([d_out, d_mask], all_rows, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, join=None):
        for _ in range(warmup):
            fn()
        if join:
            join()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        if join:
            join()  # stream-ordered: e1 fires after this rank's last strip has landed in rank 0's panorama
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if os.environ.get("ISB_BENCH_DEBUG"):
            sys.stderr.write(f"[debug] {name} rank {rank}: {float(ms.item()) / steps:.4f} ms/step over {steps} steps\n")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # join: stream-ordered wait for the runs in flight (and, with the copy-engine gather, for this rank's strips to have
    # landed in rank 0's panorama) - the closing event of the timed region is recorded behind it
    steps = args.steps if full else max(5, min(args.steps, 20))
    isb.launch_count(reset=True)
    with ClockSampler(local) as clk:
        # warm-up touches every pyramid set and both staging blocks of every slot
        ms_total = timed(step_device, steps, max(args.warmup, 2 * depth), comp.join)
    launches_per_step = isb.launch_count() // max(1, steps + max(args.warmup, 2 * depth))
    ms_step = ms_total / steps
    res = {"rig": rig, "roi": roi, "ms_step": ms_step, "value": out_mp * steps / (ms_total / 1e3), "steps": steps, "plan_s": plan_s,
           "launches_per_step": int(launches_per_step), "clocks": clk.summary(), "gather": mode, "out_mp": out_mp,
           "steps_in_flight": depth, "strip_rows": [list(r) for r in all_rows]}
    free, total = torch.cuda.mem_get_info(dev)
    res["device_mem_used_gb"] = (total - free) / 1e9  # includes the library's own cudaMalloc blocks (not torch's)

    # ---- one step at a time (pipeline depth 1): latency of a step, per-stage device time ------------------------------
    if depth > 1:
        comp.sync()
        active[0] = None
        del comp
        comp, _, _ = make_composer(1)
        active[0] = comp
        res["latency_ms_per_step"] = timed(step_device, max(5, steps // 2), 3, comp.join) / max(5, steps // 2)
    else:
        res["latency_ms_per_step"] = ms_step
    stage = {}
    for _ in range(3):
        step_device()
        for k, v in comp.timings().items():
            stage[k] = stage.get(k, 0.0) + v / 3
    comp.sync()
    res["stages_ms"] = stage

    # ---- parity ------------------------------------------------------------------------------------------------------
    # N > 1: the gathered panorama on rank 0 against the UNSHARDED result computed on rank 0 from the same inputs
    counter[0] = 0
    step_device()
    comp.sync()
    barrier()
    o8 = om = None
    parity = {}
    if rank == 0:
        d_out, d_mask = panos[0]
        if peer:
            o8 = d_out.to_numpy()[:, : pw * 3].reshape(ph, pw, 3)
            om = d_mask.to_numpy()[:, :pw]
        else:
            o8, om = d_out.cpu().numpy(), d_mask.cpu().numpy()
        if world > 1:
            c1 = isb.Composer(rig.warp, rig.scale, rig.nb, cache_plan=True)
            c1.plan(cams, sizes)
            u8 = torch.zeros((ph, pw, 3), dtype=torch.uint8, device=dev)
            um = torch.zeros((ph, pw), dtype=torch.uint8, device=dev)
            c1.run(d_imgs, d_gains, d_seams, out=u8, out_mask=um)
            torch.cuda.synchronize()
            dmax, ndiff, meq = 0, 0, True
            for r0 in range(0, ph, 2048):  # row chunks: a gigapixel panorama must not be widened to int16 in one piece
                r1 = min(ph, r0 + 2048)
                a = torch.from_numpy(np.ascontiguousarray(o8[r0:r1])).to(dev)
                dd = (a.to(torch.int16) - u8[r0:r1].to(torch.int16)).abs()
                dmax = max(dmax, int(dd.max().item()))
                ndiff += int((dd > 0).sum().item())
                meq = meq and bool(torch.equal(torch.from_numpy(np.ascontiguousarray(om[r0:r1])).to(dev), um[r0:r1]))
                del a, dd
            parity["vs_unsharded_same_rank"] = {"max_abs_diff_8bit": dmax, "n_diff": ndiff, "mask_equal": meq}
            res["byte_model"] = c1.byte_model()
            del c1, u8, um
        else:
            res["byte_model"] = comp.byte_model()
    if world > 1:
        dist.barrier()

    # ---- CPU baseline (reference CPU path, one full pass) = full-size parity vs cv2 -----------------------------------
    cpu = None
    if rank == 0 and full and not args.no_cpu_baseline and imgs is not None:
        import cv2
        cv2.ipp.setUseIPP(False)  # same setting as the reference arm; bit-exact mode of the wheel
        cv2.ocl.setUseOpenCL(False)
        tm = {}
        dt, mp, ref = cpu_reference_step(rig, imgs, gains, seams, timings=tm)
        d = np.abs(o8.astype(np.int16) - ref["result8"].astype(np.int16))
        mse = float((d.astype(np.float64) ** 2).mean())
        parity.update({"vs": f"cv2 {cv2.__version__} CPU path, full {name}", "max_abs_diff_8bit": int(d.max()),
                       "n_diff": int((d > 0).sum()), "psnr_db": 99.0 if mse == 0 else float(10 * np.log10(255 ** 2 / mse)),
                       "mask_equal": bool(np.array_equal(om, ref["mask"])),
                       "geometry_equal": bool(ref["corners"] == corners and ref["sizes"] == rsizes
                                              and tuple(ref["dst_roi"]) == tuple(roi))})
        cpu = {"value": mp / dt, "unit": UNIT, "cores": int(cv2.getNumThreads()), "kind": "reference",
               "sample": f"one full pass of {name} (all {rig.n} images) through OpenCV {cv2.__version__} "
                         f"(cv2, IPP off) in the reference's call order, {dt:.1f} s",
               "host_cpus": os.cpu_count(), "ipp": False, "stages_s": {k: round(v, 3) for k, v in tm.items()}}
        del ref, d
    if world > 1:
        dist.barrier()
    res["parity"] = parity or None
    res["cpu_baseline"] = cpu
    if rank == 0 and full and cpu is not None and world == 1 and max(pw, ph) <= 65500:
        try:
            res["output_side"] = measure_output_side(panos[0][0], panos[0][1], o8)
        except Exception as e:  # noqa: BLE001  (a side record must not take the headline line down)
            res["output_side"] = {"error": repr(e)}

    # ---- video-rate loop (BASELINE config 4): one synchronised call per frame -> latency percentiles ------------------
    if full and args.video > 0 and world == 1:
        from image_stitching_b200 import synth
        sets = [d_imgs] + [[torch.from_numpy(synth.make_image(1000 * (k + 1) + i, rig.W, rig.H)).to(dev) for i in range(rig.n)]
                           for k in range(2)]
        lat = []
        d_out, d_mask = panos[0]
        for f in range(args.video + 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            comp.run(sets[f % 3], d_gains, d_seams, out=d_out, out_mask=d_mask, out_pitch=p8, mask_pitch=pm)
            e1.record(stream)
            e1.synchronize()
            if f >= 5:
                lat.append(e0.elapsed_time(e1))
        lat = np.array(lat)
        res["video"] = {"frames": int(args.video), "p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)),
                        "max_ms": float(lat.max()), "fps": float(1e3 / lat.mean()), "MP_per_s": float(out_mp * 1e3 / lat.mean()),
                        "cached": "plan (ROIs, trig tables, tile lists); weight pyramids are rebuilt every frame"}
        del sets

    # ---- e2e: pinned host buffers through the C ABI -------------------------------------------------------------------
    if full and not args.no_e2e:
        res["e2e"] = measure_e2e(rig, imgs, gains, seams, cams, sizes, roi, args, rank, world, dev, stream, o8, om, barrier)
        if world == 1 and o8 is not None and max(roi[2], roi[3]) <= 65500 and not args.no_cpu_baseline:
            try:
                res["e2e_to_jpeg"] = measure_e2e_to_jpeg(rig, imgs, gains, seams, cams, sizes, roi, args, dev, stream, o8)
            except Exception as e:  # noqa: BLE001  (a side record must not take the headline line down)
                res["e2e_to_jpeg"] = {"error": repr(e)}
    del d_imgs, panos
    return res


def measure_e2e(rig, imgs, gains, seams, cams, sizes, roi, args, rank, world, dev, stream, o8_ref, om_ref, barrier):
    """Host -> host through the C ABI.  `depth` composers in async mode on `depth` streams per rank (a double/triple-buffered
    serving loop): step k's upload overlaps step k-1's download on the full-duplex PCIe link.  Every step uploads its inputs
    from pinned host memory and downloads its panorama rows into pinned host memory."""
    import torch
    import torch.distributed as dist

    import image_stitching_b200 as isb
    pw, ph = roi[2], roi[3]
    depth = max(1, args.e2e_depth)
    src_bytes = rig.n * rig.W * rig.H * 3
    out_bytes = ph * pw * 4
    shared = None
    if world > 1:
        try:
            shared = SharedPinned(src_bytes + depth * out_bytes, rank, world, "e2e")
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"rank {rank}: shared pinned host memory unavailable ({e})\n")
            shared = None
        ok = torch.tensor([1 if shared is not None else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not bool(ok.item()):
            if shared is not None:
                shared.close()
            return {"unavailable": "shared pinned host memory could not be set up on this node"}
    if shared is not None:
        h_imgs = [shared.view(i * rig.W * rig.H * 3, (rig.H, rig.W, 3)) for i in range(rig.n)]
        if rank == 0:
            for d, s in zip(h_imgs, imgs):
                d[...] = s
        outs = [(shared.view(src_bytes + k * out_bytes, (ph, pw, 3)), shared.view(src_bytes + k * out_bytes + ph * pw * 3, (ph, pw)))
                for k in range(depth)]
        dist.barrier()
    else:
        h_imgs = [torch.from_numpy(im).pin_memory().numpy() for im in imgs]
        outs = [(torch.zeros((ph, pw, 3), dtype=torch.uint8).pin_memory().numpy(), torch.zeros((ph, pw), dtype=torch.uint8).pin_memory().numpy())
                for _ in range(depth)]
    h_gains = [torch.from_numpy(g).pin_memory().numpy() for g in gains]
    h_seams = [torch.from_numpy(s_).pin_memory().numpy() for s_ in seams]
    streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
    comps = []
    for _ in range(depth):
        c2 = isb.Composer(rig.warp, rig.scale, rig.nb, strip_index=rank, strip_count=world, cache_plan=True, async_mode=True)
        c2.plan(cams, sizes)
        comps.append(c2)
    torch.cuda.synchronize()

    def pipelined(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s_ in streams:
            s_.wait_event(e0)
        for k in range(steps):
            isb.set_stream(streams[k % depth].cuda_stream)
            comps[k % depth].run(h_imgs, h_gains, h_seams, out=outs[k % depth][0], out_mask=outs[k % depth][1])
        for s_ in streams:
            stream.wait_stream(s_)
        e1.record(stream)
        barrier()
        isb.set_stream(stream.cuda_stream)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def sync_one(steps):  # depth 1: one synchronous call per step, H2D -> kernels -> D2H back to back
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            comps[0].run(h_imgs, h_gains, h_seams, out=outs[0][0], out_mask=outs[0][1])
            comps[0].sync()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    e2e_steps = max(2, min(args.steps, 50))
    isb.set_stream(stream.cuda_stream)
    sync_one(2)
    ms_sync = sync_one(max(2, e2e_steps // 4))
    sync_steps = max(2, e2e_steps // 4)
    pipelined(2 * depth)
    runs_ms = [pipelined(e2e_steps) for _ in range(2)]  # the host side of a shared box is noisy: both runs are reported, the better one counts
    ms_pipe = min(runs_ms)
    h2d = comps[0].last_h2d_bytes() + sum(g.nbytes for g in gains) + sum(s.nbytes for s in seams)
    rows = comps[0].strip_rows
    d2h = (rows[1] - rows[0]) * pw * 4
    tot = torch.tensor([h2d, d2h], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    equal = None
    if rank == 0 and o8_ref is not None:
        equal = bool(all(np.array_equal(o[0], o8_ref) and np.array_equal(o[1], om_ref) for o in outs))
    out_mp = pw * ph / 1e6
    r = {"value": out_mp * e2e_steps / (ms_pipe / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
         "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": ms_pipe / e2e_steps, "steps": e2e_steps,
         "runs_ms_per_step": [m / e2e_steps for m in runs_ms], "sync_ms_per_step": ms_sync / sync_steps, "pipeline_depth": depth, "host_result_equals_device_result": equal,
         "host_memory": "pinned, shared by the node's ranks (one copy of sources and panorama)" if shared is not None else "pinned",
         "bytes_are": "summed over ranks: each rank uploads the source row bands its strip reads and downloads its strip"}
    del comps
    if shared is not None:
        barrier()
        shared.close()
    return r


def measure_e2e_to_jpeg(rig, imgs, gains, seams, cams, sizes, roi, args, dev, stream, o8_ref):
    """The reference's whole deliverable, frames in -> result.jpg out (image_stitching.cpp:1086-1229 incl. the imwrite of :1228),
    end to end on one GPU: every step uploads the decoded frames from pinned host memory, composes the panorama on the device,
    encodes it there (isb_jpeg_encode) and downloads the FILE (tens of MB) instead of the raw panorama (241 MB for cfg2).  The
    encoder returns when its file is in host memory, so `depth` host threads - one composer, stream and output slot each - keep
    the PCIe link busy the way the pipelined composers of `e2e` do."""
    import ctypes as C
    import threading as th

    import cv2
    import torch

    import image_stitching_b200 as isb
    pw, ph = roi[2], roi[3]
    depth = max(1, args.e2e_depth)
    steps = max(depth, min(args.steps, 50))
    h_imgs = [torch.from_numpy(im).pin_memory().numpy() for im in imgs]
    h_gains = [torch.from_numpy(g).pin_memory().numpy() for g in gains]
    h_seams = [torch.from_numpy(s_).pin_memory().numpy() for s_ in seams]
    cap = pw * ph
    slots = []
    for _ in range(depth):
        c = isb.Composer(rig.warp, rig.scale, rig.nb, cache_plan=True, async_mode=True)
        c.plan(cams, sizes)
        slots.append(dict(comp=c, stream=torch.cuda.Stream(device=dev), d8=torch.zeros((ph, pw, 3), dtype=torch.uint8, device=dev),
                          dm=torch.zeros((ph, pw), dtype=torch.uint8, device=dev), file=torch.zeros(cap, dtype=torch.uint8).pin_memory(),
                          n=C.c_size_t(0), err=None))
    L = isb.lib()

    # persistent workers (the encoder's work buffers live per thread): phases = warm-up, run 1, run 2
    phases = [2 * depth, steps, steps]
    gate = th.Barrier(depth + 1)

    def worker(k):
        sl = slots[k]
        torch.cuda.set_device(dev)
        isb.set_stream(sl["stream"].cuda_stream)
        for n_total in phases:
            n_steps = n_total // depth + (1 if k < n_total % depth else 0)
            gate.wait()  # start of the phase
            try:
                for _ in range(n_steps if not sl["err"] else 0):
                    sl["comp"].run(h_imgs, h_gains, h_seams, out=sl["d8"], out_mask=sl["dm"])
                    rc = L.isb_jpeg_encode(C.c_void_p(sl["d8"].data_ptr()), pw, ph, C.c_size_t(pw * 3), 0, 95, C.c_void_p(sl["file"].data_ptr()),
                                           C.c_size_t(cap), C.byref(sl["n"]))
                    if rc != 0:
                        raise RuntimeError(f"isb_jpeg_encode failed ({rc})")
            except Exception as e:  # noqa: BLE001
                sl["err"] = repr(e)
            gate.wait()  # end of the phase

    ts = [th.Thread(target=worker, args=(k,)) for k in range(depth)]
    for t in ts:
        t.start()

    def loop():
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for sl in slots:
            sl["stream"].wait_event(e0)
        t0 = time.perf_counter()
        gate.wait()
        gate.wait()
        wall = (time.perf_counter() - t0) * 1e3
        for sl in slots:
            stream.wait_stream(sl["stream"])
        e1.record(stream)
        torch.cuda.synchronize()
        return max(e0.elapsed_time(e1), wall)  # the files are in host memory when the workers return: the later clock counts

    loop()
    runs = [loop() for _ in range(2)]
    for t in ts:
        t.join()
    for sl in slots:
        if sl["err"]:
            raise RuntimeError(sl["err"])
    ms = min(runs) / steps
    ok, ref = cv2.imencode(".jpg", o8_ref)
    refb = ref.tobytes()
    equal = bool(ok and all(sl["file"][: sl["n"].value].numpy().tobytes() == refb for sl in slots))
    isb.set_stream(stream.cuda_stream)
    out_mp = pw * ph / 1e6
    h2d = slots[0]["comp"].last_h2d_bytes() + sum(g.nbytes for g in gains) + sum(s_.nbytes for s_ in seams)
    r = {"value": out_mp / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "runs_ms_per_step": [m / steps for m in runs],
         "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(slots[0]["n"].value), "host_threads": depth,
         "file_equals_cv2_imencode_of_the_panorama": equal,
         "what": "decoded frames in pinned host memory -> result.jpg bytes in pinned host memory (compose + imwrite's JPEG encoding on the device)"}
    del slots
    return r


def measure_output_side(d_out, d_mask, o8):
    """What follows blend() in the reference: imwrite(result_name, result) (image_stitching.cpp:1228) and the crop() of cropper.cpp
    on the composited mask.  GPU: isb_jpeg_encode from the device-resident panorama to the file bytes in host memory, and
    isb_crop_rect on the device-resident mask; CPU: cv2.imencode of the same panorama (libjpeg-turbo, what imwrite calls)."""
    import cv2
    import torch

    import image_stitching_b200 as isb
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        jpg = isb.imencode_jpg(d_out)
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ok, ref = cv2.imencode(".jpg", o8)
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rect = isb.crop_rect(d_mask)
    t_crop = time.perf_counter() - t0
    isb.lib().isb_jpeg_release_workspace()
    mp = o8.shape[0] * o8.shape[1] / 1e6
    return {"imwrite_jpg": {"gpu_ms": min(ts) * 1e3, "gpu_MP_per_s": mp / min(ts), "cpu_ms": t_cpu * 1e3, "cpu_MP_per_s": mp / t_cpu,
                            "bytes": len(jpg), "equal_to_cv2_imencode": bool(ok and jpg == ref.tobytes()),
                            "note": "quality 95, 4:2:0, baseline - imwrite's defaults; GPU time includes the copy of the file bytes to host memory"},
            "crop": {"rect_xywh": [int(v) for v in rect], "gpu_ms": t_crop * 1e3}}


def measure_cfg1_substitute(args, dev, stream):
    """BASELINE config 0 stand-in (samples.zip is a git-lfs pointer): the reference's DEFAULT flow - rotate(180), compose_megapix
    = 0.4 INTER_LINEAR_EXACT resize, seam masks at seam_megapix = 0.1, the band count of image_stitching.cpp:1183 - on the
    full-size cfg2 frames.  CPU: OpenCV through cv2 in the reference's call order, timed; GPU: ONE isb_composer_run() per step
    from the decoded frames in pinned host memory (ingest pre-steps inside the composer) to the panorama in host memory."""
    import cv2
    import torch

    import image_stitching_b200 as isb
    from image_stitching_b200 import synth
    from oracle import cv_reference as cvr
    rig, imgs, gains = make_inputs("cfg2", 1)
    fs = synth.default_flow_setup(rig)
    src_mask = synth.seam_band_mask(*fs["seam_size"])
    cv2.ipp.setUseIPP(False)
    cv2.ocl.setUseOpenCL(False)
    # --- CPU reference: seam-scale warps + rotate + resize + the loop
    t0 = time.perf_counter()
    seams = [cvr.make_warper(rig.warp, fs["seam_warper_scale"]).warp(src_mask, K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)[1]
             for K, R in zip(fs["Ks"], rig.Rs)]
    t_seam = time.perf_counter() - t0
    w = cvr.make_warper(rig.warp, fs["scale_c"])
    rois = [w.warpRoi(fs["sz"], K, R) for K, R in zip(fs["Kc"], rig.Rs)]
    roi = cv2.detail.resultRoi(corners=[(r[0], r[1]) for r in rois], sizes=[(r[2], r[3]) for r in rois])
    nb = isb.num_bands_for(roi[2], roi[3], 5.0)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        small = [cv2.resize(cv2.rotate(im, cv2.ROTATE_180), None, fx=fs["compose_scale"], fy=fs["compose_scale"],
                            interpolation=cv2.INTER_LINEAR_EXACT) for im in imgs]
        ref = cvr.compose_cv(small, fs["Kc"], rig.Rs, fs["scale_c"], rig.warp, nb, gains, seams)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out_mp = roi[2] * roi[3] / 1e6
    # --- GPU: decoded frames (pinned host) -> panorama (pinned host), one call per step
    comp = isb.Composer(rig.warp, fs["scale_c"], 0, ingest_rotate=isb.ROTATE_180, compose_scale=fs["compose_scale"],
                        blend_type="multiband", blend_strength=5.0, cache_plan=True)
    cams = isb.cameras_from_KR(fs["Kc"], rig.Rs)
    _, _, groi = comp.plan(cams, [(rig.W, rig.H)] * rig.n)
    h_imgs = [torch.from_numpy(im).pin_memory().numpy() for im in imgs]
    ho = torch.zeros((groi[3], groi[2], 3), dtype=torch.uint8).pin_memory().numpy()
    hm = torch.zeros((groi[3], groi[2]), dtype=torch.uint8).pin_memory().numpy()
    for _ in range(3):
        comp.run(h_imgs, gains, seams, out=ho, out_mask=hm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 10
    e0.record(stream)
    for _ in range(k):
        comp.run(h_imgs, gains, seams, out=ho, out_mask=hm)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / k
    d = np.abs(ho.astype(np.int16) - ref["result8"].astype(np.int16))
    return {"what": "reference default flow (compose_megapix 0.4, seam_megapix 0.1, blend_strength 5) on the 8 x 12 MP cfg2 frames: "
                    "rotate(180) + INTER_LINEAR_EXACT resize + warp + multi-band blend",
            "panorama": [int(roi[2]), int(roi[3])], "output_MP": out_mp, "num_bands": int(nb), "compose_scale": fs["compose_scale"],
            "cpu_reference": {"value": out_mp / best, "unit": UNIT, "seconds": best, "seam_scale_warps_s": t_seam,
                              "cores": int(cv2.getNumThreads()), "kind": "reference",
                              "sample": f"all {rig.n} frames, best of 3 passes, OpenCV {cv2.__version__} (cv2, IPP off)"},
            "ours_e2e": {"value": out_mp / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
                         "h2d_bytes_per_step": int(sum(im.nbytes for im in imgs)), "d2h_bytes_per_step": int(ho.nbytes + hm.nbytes),
                         "note": "one isb_composer_run() per step from the decoded 12 MP frames in pinned host memory"},
            "parity": {"max_abs_diff_8bit": int(d.max()), "n_diff": int((d > 0).sum()), "mask_equal": bool(np.array_equal(hm, ref["mask"])),
                       "geometry_equal": bool(tuple(groi) == tuple(int(v) for v in roi))}}


def roofline_of(res, world, workload):
    peak, peak_src = peaks()
    bm = res["byte_model"]
    rig = res["rig"]
    P = sum(4.0 ** -l for l in range(rig.nb + 1))
    S, M, Ap = bm["S"], bm["M"], bm["Ap"]
    alg = {"warp": 3 * S + 10 * M, "pyrdown": 10 * M * (P - 1), "blend": 30 * P * M + 10 * P * Ap + 4 * Ap}
    ms_step, stage = res["ms_step"], res["stages_ms"]
    achieved = bm["B_alg"] / (ms_step / 1e3) / 1e9
    # measured DRAM traffic (ncu dram__bytes_read + write over the step's kernels) exists for the profiled workload on one GPU
    traffic = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(tp):
        try:
            t = json.load(open(tp))
            ent = t.get("workloads", {}).get(workload)
            if ent:
                traffic, traffic_src = ent.get("dram_bytes_per_step"), ent.get("source")
        except Exception:
            pass
    r = {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
         "peak_per_gpu": peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
         "kernel": "whole compose step: seam_prep + warp_tiles_packed + pyrdown x nb + blend x (nb+1)",
         "algorithmic_bytes_per_step": bm["B_alg"], "S_px": S, "M_px": M, "Ap_px": Ap,
         "note": "frac = B_alg (SURVEY 8(d): includes the reference's accumulator read-modify-write, which this output-centric "
                 "design never performs) / step time / peak; frac_dram = bytes actually moved (ncu) / step time / peak",
         "stages_ms": stage,
         "stages_frac": {k: (alg[k] / (stage[k] / 1e3) / 1e9 / (peak * world) if stage.get(k) else None) for k in alg}}
    if traffic:
        r["achieved_dram"] = traffic / (ms_step / 1e3) / 1e9
        r["frac_dram"] = r["achieved_dram"] / peak
    return r


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist

    import image_stitching_b200 as isb

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    isb.set_stream(stream.cuda_stream)
    args.steps = args.steps if args.steps is not None else 50

    res = measure(args.workload, args.div, args, rank, world, dev, stream, full=True)
    sub = None
    if args.workload == "cfg2" and args.div == 1 and not args.no_cfg3:
        # the 36 x 24 MP rig the north star names for strip scaling, in the same line at every N (device-resident value)
        torch.cuda.empty_cache()
        sub = measure("cfg3", 1, args, rank, world, dev, stream, full=False)
    cfg1 = None
    if rank == 0 and world == 1 and args.workload == "cfg2" and args.div == 1 and not args.no_cfg1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        try:
            cfg1 = measure_cfg1_substitute(args, dev, stream)
        except Exception as e:  # noqa: BLE001  (a side record must not take the headline line down)
            cfg1 = {"error": repr(e)}
    if rank == 0:
        rig, roi = res["rig"], res["roi"]
        par = f"strips{world}" + ("" if world == 1 else f"+{res['gather']}-gather")
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8/s16/f32", "data": "synthetic", "config": config_of(rig, args.workload, args.div, roi),
                "latency_ms_per_step": res["latency_ms_per_step"],
                "layout": {"parallelism": par, "plan_cache": True, "plan_s": res["plan_s"], "steps_in_flight": res["steps_in_flight"],
                           "strip_rows": res["strip_rows"],
                           "steps_in_flight_note": "ms_per_step / value are the throughput of back-to-back steps with that many "
                                                   "runs in flight (own pyramids and panorama each); latency_ms_per_step is one "
                                                   "step at a time",
                           "l2": "per-step working set (sources + per-image pyramids) >> 126 MB L2; no explicit flush",
                           "device_mem_used_gb_rank0": res["device_mem_used_gb"]},
                "e2e": res.get("e2e"), "gpu_launches": int(res["launches_per_step"] * args.steps),
                "gpu_launches_per_step": res["launches_per_step"], "clocks": res["clocks"],
                "roofline": roofline_of(res, world, args.workload), "cpu_baseline": res["cpu_baseline"], "parity": res["parity"]}
        if "video" in res:
            line["video"] = res["video"]
        if sub is not None:
            rl = roofline_of(sub, world, "cfg3")
            line["cfg3"] = {"config": config_of(sub["rig"], "cfg3", 1, sub["roi"]), "value": sub["value"], "unit": UNIT,
                            "ms_per_step": sub["ms_step"], "latency_ms_per_step": sub["latency_ms_per_step"], "steps": sub["steps"],
                            "steps_in_flight": sub["steps_in_flight"], "strip_rows": sub["strip_rows"], "parallelism": f"strips{world}" + ("" if world == 1 else f"+{sub['gather']}-gather"),
                            "roofline_frac": rl["frac"], "algorithmic_bytes_per_step": rl["algorithmic_bytes_per_step"],
                            "stages_ms": sub["stages_ms"], "parity": sub["parity"], "plan_s": sub["plan_s"],
                            "device_mem_used_gb_rank0": sub["device_mem_used_gb"],
                            "data": "synthetic, generated on the device (same construction as the numpy rig)"}
        if cfg1 is not None:
            line["cfg1_substitute"] = cfg1
        if res.get("output_side"):
            line["output_side"] = res["output_side"]
        if res.get("e2e_to_jpeg"):
            line["e2e_to_jpeg"] = res["e2e_to_jpeg"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 50; reference arm: 10)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--div", type=int, default=1, help="linear down-scale of the rig (dev only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg3 sub-record of the default (cfg2) line")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--device-images", action="store_true",
                    help="generate the workload's images on the device (for rigs whose numpy generation takes minutes, e.g. cfg5 at "
                         "full size); implies no CPU baseline and no e2e, parity is then strips vs unsharded only")
    ap.add_argument("--no-cfg1", action="store_true", help="skip the cfg1-substitute (reference default flow) sub-record")
    ap.add_argument("--inflight", type=int, default=3, help="runs in flight in the device-resident throughput loop (isb_config.pipeline_depth)")
    ap.add_argument("--video", type=int, default=0, help="also run N frames one call at a time and report p50/p95 latency")
    ap.add_argument("--e2e-depth", type=int, default=3, help="composers in flight per rank in the pipelined e2e measurement")
    ap.add_argument("--gather", default="copy", choices=["copy", "p2p", "nccl"],
                    help="N > 1: strips reach rank 0's panorama by the copy engine from a local staging block (copy), by peer "
                         "stores of the final kernel (p2p), or by NCCL send/recv (nccl)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.device_images:
        args.no_cpu_baseline = args.no_e2e = True
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()

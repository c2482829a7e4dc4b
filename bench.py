#!/usr/bin/env python
"""bench.py - output-panorama MP/s of the fused warp + multi-band blend on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

One "step" = one pass of the compositing hot path (image_stitching.cpp:1086-1229: warp, mask, gain, ->16S,
seam mask, MultiBandBlender prepare/feed/blend, saturate to 8U) over the synthetic cfg2 rig of SURVEY.md 8(d):
8 x 4000x3000 images, spherical warp, 5 bands, 20912x2881 panorama.

  value  : device-resident inputs -> device-resident panorama (on rank 0 when N > 1), CUDA-event timed.
  e2e    : the same step through the C ABI with pinned HOST buffers (H2D of the sources and D2H of the
           panorama inside the timed region).
  N > 1  : the panorama is cut into N horizontal strips (2^nb-aligned, 4*2^nb halo rows recomputed); each
           rank composes its strip and the strips are gathered on rank 0 over NCCL.  scaling = "strong".
  --impl reference : the reference's own CPU implementation of the path (OpenCV's cv::detail classes, driven
           through cv2 in the reference's call order) on the host cores; rank 0 only.
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


def make_inputs(workload, div=1):
    from image_stitching_b200 import synth
    rig = synth.make_rig(workload, scale_div=div)
    imgs = [synth.make_image(i, rig.W, rig.H) for i in range(rig.n)]
    gains = synth.make_gains(rig.n)
    return rig, imgs, gains


def synth_image(index, rig):
    from image_stitching_b200 import synth
    return synth.make_image(index, rig.W, rig.H)


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


def cpu_reference_step(rig, imgs, gains, seams, n_images=None, timings=None):
    """One pass of the reference CPU path on the first n_images of the rig; returns (seconds, output MP, result)."""
    from oracle import cv_reference as cvr
    n = rig.n if n_images is None else n_images
    t0 = time.perf_counter()
    ref = cvr.compose_cv(imgs[:n], rig.Ks[:n], rig.Rs[:n], rig.scale, rig.warp, rig.nb, gains[:n],
                         None if seams is None else seams[:n], timings=timings)
    dt = time.perf_counter() - t0
    return dt, ref["dst_roi"][2] * ref["dst_roi"][3] / 1e6, ref


def seam_masks_cpu(rig):
    from oracle import cv_reference as cvr
    return cvr.seam_masks_cv(rig.warp, rig.scale, rig.Ks, rig.Rs, rig.W, rig.H)


def run_reference(args, rank, world):
    if rank != 0:
        return
    import cv2
    from oracle import cv_reference as cvr
    cv2.ipp.setUseIPP(True)  # timing: the wheel as shipped
    cv2.ocl.setUseOpenCL(False)
    rig, imgs, gains = make_inputs(args.workload, args.div)
    seams = seam_masks_cpu(rig)
    # bounded sample: as many leading images of the rig as fit a ~150 s budget for the whole K+W run
    t1, _, _ = cpu_reference_step(rig, imgs, gains, seams, n_images=1)
    budget = 150.0 / max(1, args.steps + args.warmup)
    n_s = int(max(1, min(rig.n, budget // max(t1, 1e-3))))
    for _ in range(args.warmup):
        cpu_reference_step(rig, imgs, gains, seams, n_images=n_s)
    tot, mp = 0.0, 0.0
    for _ in range(args.steps):
        dt, m, _ = cpu_reference_step(rig, imgs, gains, seams, n_images=n_s)
        tot += dt
        mp += m
    val = mp / tot
    cores = cv2.getNumThreads()
    sample = (f"first {n_s} of {rig.n} images of {args.workload} ({rig.W}x{rig.H}, {rig.warp}, {rig.nb} bands) per step; "
              f"OpenCV {cv2.__version__} (cv2 wheel, IPP on) cv::detail warper/compensator/MultiBandBlender in the "
              f"reference's call order; the reference's main() needs OpenCV dev files + libexif and cannot be built here")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8/s16/f32", "data": "synthetic",
            "config": {"workload": workload_name(rig, args)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": int(cores), "kind": "reference", "sample": sample,
                             "host_cpus": os.cpu_count()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(rig, args):
    return (f"{args.workload}{'' if args.div == 1 else '/div' + str(args.div)}: {rig.n}x{rig.W}x{rig.H} 8UC3, {rig.warp} warp, "
            f"{rig.nb}-band multi-band blend")


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

    rig, imgs, gains = make_inputs(args.workload, args.div)
    seams = seam_masks_gpu(rig)
    comp = isb.Composer(rig.warp, rig.scale, rig.nb, strip_index=rank, strip_count=world, cache_plan=True)
    cams = isb.cameras_from_KR(rig.Ks, rig.Rs)
    sizes = [(rig.W, rig.H)] * rig.n
    corners, rsizes, roi = comp.plan(cams, sizes)
    pw, ph = roi[2], roi[3]
    out_mp = pw * ph / 1e6

    # device-resident inputs and outputs
    d_imgs = [torch.from_numpy(im).to(dev) for im in imgs]
    d_gains = [torch.from_numpy(g).to(dev) for g in gains]
    d_seams = [torch.from_numpy(s).to(dev) for s in seams]
    from image_stitching_b200 import strips
    m_ = 1 << rig.nb  # the rigs' band counts are below the prepare() clamp, so nb is the actual band count
    padded_h = (ph + m_ - 1) // m_ * m_
    all_rows = strips.all_strip_rows(padded_h, ph, rig.nb, world)
    use_p2p = world > 1 and args.gather == "p2p"
    # peer-memory gather: rows padded to 128 B so that the level-0 blend kernel's staged 16-byte stores apply over NVLink
    p8, pm = (pw * 3, pw) if not use_p2p else ((pw * 3 + 127) // 128 * 128, (pw + 127) // 128 * 128)
    if use_p2p:
        # fused collapse + gather: rank 0 owns the panorama, the other ranks map it through CUDA IPC and their final
        # blend kernel stores its rows straight into rank 0's HBM over NVLink (no separate collective, no staging)
        ok = 1
        try:
            if rank == 0:
                d_out, d_mask = isb.DevPtr.alloc((ph, p8)), isb.DevPtr.alloc((ph, pm))
                handles = [d_out.ipc_handle(), d_mask.ipc_handle()]
            else:
                handles = [None, None]
        except Exception as e:  # noqa: BLE001
            handles, ok = [None, None], 0
            sys.stderr.write(f"rank {rank}: peer-memory export failed ({e}); falling back to NCCL gather\n")
        dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            try:
                d_out, d_mask = isb.DevPtr.open_ipc(handles[0], (ph, p8)), isb.DevPtr.open_ipc(handles[1], (ph, pm))
            except Exception as e:  # noqa: BLE001
                ok = 0
                sys.stderr.write(f"rank {rank}: peer-memory import failed ({e}); falling back to NCCL gather\n")
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        use_p2p = bool(flag.item())  # every rank takes the same path
    if not use_p2p:
        p8, pm = pw * 3, pw
        d_out = torch.zeros((ph, pw, 3), dtype=torch.uint8, device=dev)
        d_mask = torch.zeros((ph, pw), dtype=torch.uint8, device=dev)

    def gather_strips(rows):
        # strips -> rank 0 over NCCL (full-width rows are contiguous slices of the panorama tensors)
        assert tuple(rows) == tuple(all_rows[rank])
        if not use_p2p:
            strips.gather_strips([d_out, d_mask], all_rows, rank, world)

    def step_device():
        r = comp.run(d_imgs, d_gains, d_seams, out=d_out, out_mask=d_mask, out_pitch=p8, mask_pitch=pm)
        if world > 1:
            gather_strips(r["strip_rows"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value: device-resident -------------------------------------------------------------------
    isb.launch_count(reset=True)
    with ClockSampler(local) as clk:
        ms_total = timed(step_device, args.steps, args.warmup)
    launches_per_step = isb.launch_count() // max(1, args.steps + args.warmup)
    ms_step = ms_total / args.steps
    value = out_mp * args.steps / (ms_total / 1e3)

    video = None
    if args.video > 0 and world == 1:
        # BASELINE config 4 style serving loop: fixed cameras / masks / gains (plan cached), new pixels every frame
        # (three resident frame sets cycled), one synchronised call per frame -> latency percentiles + throughput
        sets = [d_imgs] + [[torch.from_numpy(synth_image(1000 * (k + 1) + i, rig)).to(dev) for i in range(rig.n)] for k in range(2)]
        lat = []
        for f in range(args.video + 5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            comp.run(sets[f % 3], d_gains, d_seams, out=d_out, out_mask=d_mask, out_pitch=p8, mask_pitch=pm)
            e1.record(stream)
            e1.synchronize()
            if f >= 5:
                lat.append(e0.elapsed_time(e1))
        lat = np.array(lat)
        video = {"frames": int(args.video), "p50_ms": float(np.percentile(lat, 50)), "p95_ms": float(np.percentile(lat, 95)),
                 "max_ms": float(lat.max()), "fps": float(1e3 / lat.mean()), "MP_per_s": float(out_mp * 1e3 / lat.mean()),
                 "cached": "plan (ROIs, trig tables, tile lists); weight pyramids are rebuilt every frame"}

    # per-stage device time (separate short loop so the event syncs do not perturb the timed region)
    stage = {}
    for _ in range(3):
        comp.run(d_imgs, d_gains, d_seams, out=d_out, out_mask=d_mask, out_pitch=p8, mask_pitch=pm)
        for k, v in comp.timings().items():
            stage[k] = stage.get(k, 0.0) + v / 3

    # ---- e2e: pinned host buffers through the C ABI ---------------------------------------------------
    h_imgs = [torch.from_numpy(im).pin_memory() for im in imgs]
    # (peer-memory gather: the host copies keep the padded row pitch of rank 0's panorama)
    h_out = torch.zeros((ph, pw, 3) if p8 == pw * 3 else (ph, p8), dtype=torch.uint8).pin_memory()
    h_mask = torch.zeros((ph, pm), dtype=torch.uint8).pin_memory()
    h2d = sum(int(t.numel()) for t in h_imgs) + sum(g.nbytes for g in gains) + sum(s.nbytes for s in seams)
    d2h = int(h_out.numel() + h_mask.numel()) if rank == 0 else 0

    def step_e2e():
        if world == 1:
            comp.run([t.numpy() for t in h_imgs], gains, seams, out=h_out.numpy(), out_mask=h_mask.numpy())
        else:
            r = comp.run([t.numpy() for t in h_imgs], gains, seams, out=d_out, out_mask=d_mask, out_pitch=p8, mask_pitch=pm)
            gather_strips(r["strip_rows"])
            if use_p2p:
                torch.cuda.synchronize()
                dist.barrier()  # every rank's rows have landed in rank 0's panorama
            if rank == 0:
                if use_p2p:
                    d_out.to_numpy(out=h_out.numpy())
                    d_mask.to_numpy(out=h_mask.numpy())
                else:
                    h_out.copy_(d_out, non_blocking=True)
                    h_mask.copy_(d_mask, non_blocking=True)
                    torch.cuda.synchronize()

    e2e_steps = max(2, min(args.steps, 50))
    ms_e2e_sync = timed(step_e2e, e2e_steps, 2)  # one synchronous call per step: H2D -> kernels -> D2H back to back
    e2e_extra = {"sync_ms_per_step": ms_e2e_sync / e2e_steps, "pipeline_depth": 1}
    ms_e2e = ms_e2e_sync
    if world == 1 and not args.no_e2e_pipeline:
        # Same call, two composers in async_mode on two streams (a double-buffered serving loop, e.g. video-rate cfg4):
        # step k's upload overlaps step k-1's download on the full-duplex PCIe link.  Every step still uploads its
        # inputs from pinned host memory and downloads its panorama + mask.
        depth = args.e2e_depth
        streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        slots = []
        for _ in range(depth):
            c2 = isb.Composer(rig.warp, rig.scale, rig.nb, cache_plan=True, async_mode=True)
            c2.plan(cams, sizes)
            slots.append((c2, torch.zeros((ph, pw, 3), dtype=torch.uint8).pin_memory(),
                          torch.zeros((ph, pw), dtype=torch.uint8).pin_memory()))
        h_gains = [torch.from_numpy(g).pin_memory() for g in gains]
        h_seams = [torch.from_numpy(s_).pin_memory() for s_ in seams]
        np_imgs = [t.numpy() for t in h_imgs]
        np_gains, np_seams = [t.numpy() for t in h_gains], [t.numpy() for t in h_seams]
        torch.cuda.synchronize()

        def pipelined(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for s_ in streams:
                s_.wait_event(e0)
            for k in range(steps):
                c2, ho, hm = slots[k % depth]
                isb.set_stream(streams[k % depth].cuda_stream)
                c2.run(np_imgs, np_gains, np_seams, out=ho.numpy(), out_mask=hm.numpy())
            for s_ in streams:
                stream.wait_stream(s_)
            e1.record(stream)
            torch.cuda.synchronize()
            isb.set_stream(stream.cuda_stream)
            return e0.elapsed_time(e1)

        pipelined(2 * depth)
        ms_e2e = pipelined(e2e_steps)
        # the async path must give the same panorama as the device-resident one
        comp.run(d_imgs, d_gains, d_seams, out=d_out, out_mask=d_mask, out_pitch=p8, mask_pitch=pm)
        torch.cuda.synchronize()
        e2e_extra.update({"pipeline_depth": depth,
                          "pipelined_equals_device_result": bool(torch.equal(slots[0][1], d_out.cpu()) and
                                                                 torch.equal(slots[-1][2], d_mask.cpu()))})
        del slots
    e2e_value = out_mp * e2e_steps / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline (SURVEY.md 8(d) byte model, whole step) ---------------------------------------------
    peak, peak_src = peaks()
    bm = comp.byte_model() if world == 1 else None
    if bm is None:
        c1 = isb.Composer(rig.warp, rig.scale, rig.nb)
        c1.plan(cams, sizes)
        bm = c1.byte_model()
    P = sum(4.0 ** -l for l in range(rig.nb + 1))
    S, M, Ap = bm["S"], bm["M"], bm["Ap"]
    alg = {"warp": 3 * S + 10 * M, "pyrdown": 10 * M * (P - 1), "blend": 30 * P * M + 10 * P * Ap + 4 * Ap}
    achieved = bm["B_alg"] / (ms_step / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_step")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                "peak_per_gpu": peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "whole compose step: warp_tiles + pyrdown_tiles x nb + blend_level x (nb+1)",
                "algorithmic_bytes_per_step": bm["B_alg"], "S_px": S, "M_px": M, "Ap_px": Ap,
                "stages_ms": stage,
                "stages_frac": {k: (alg[k] / (stage[k] / 1e3) / 1e9 / (peak * world) if stage.get(k) else None) for k in alg}}

    # ---- CPU baseline (reference CPU path, bounded: one full pass) + full-size parity --------------------
    cpu = None
    parity = None
    if not args.no_cpu_baseline and world == 1:
        import cv2
        cv2.ipp.setUseIPP(False)  # this pass doubles as the full-size parity check (parity mode)
        cv2.ocl.setUseOpenCL(False)
        tm = {}
        dt, mp, ref = cpu_reference_step(rig, imgs, gains, seams, timings=tm)
        ours = comp.run(d_imgs, d_gains, d_seams, out=d_out, out_mask=d_mask, out_pitch=p8, mask_pitch=pm)
        torch.cuda.synchronize()
        o8, om = d_out.cpu().numpy(), d_mask.cpu().numpy()
        d = np.abs(o8.astype(np.int16) - ref["result8"].astype(np.int16))
        mse = float((d.astype(np.float64) ** 2).mean())
        parity = {"vs": f"cv2 {cv2.__version__} CPU path, full {args.workload}", "max_abs_diff_8bit": int(d.max()),
                  "n_diff": int((d > 0).sum()), "psnr_db": 99.0 if mse == 0 else float(10 * np.log10(255 ** 2 / mse)),
                  "mask_equal": bool(np.array_equal(om, ref["mask"])),
                  "geometry_equal": bool(ref["corners"] == ours["corners"] and ref["sizes"] == ours["sizes"]
                                         and tuple(ref["dst_roi"]) == tuple(ours["dst_roi"]))}
        cpu = {"value": mp / dt, "unit": UNIT, "cores": int(cv2.getNumThreads()), "kind": "reference",
               "sample": f"one full pass of {args.workload} ({rig.n} images) through OpenCV {cv2.__version__} "
                         f"(cv2, IPP off = parity mode) in the reference's call order, {dt:.1f} s",
               "host_cpus": os.cpu_count(), "stages_s": {k: round(v, 3) for k, v in tm.items()}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8/s16/f32", "data": "synthetic",
            "config": {"workload": workload_name(rig, args), "panorama": [pw, ph], "output_MP": out_mp,
                       "parallelism": f"strips{world}" + ("" if world == 1 else ("+p2p" if use_p2p else "+nccl") + "-gather"), "plan_cache": True,
                       "l2": "per-step working set (sources + per-image pyramids) >> 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, **e2e_extra},
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
            "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu, "parity": parity}
    if video is not None:
        line["video"] = video
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--div", type=int, default=1, help="linear down-scale of the rig (dev only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--video", type=int, default=0, help="also run N frames one call at a time and report p50/p95 latency")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="report the synchronous one-call-per-step e2e only")
    ap.add_argument("--e2e-depth", type=int, default=3, help="composers in flight in the pipelined e2e measurement")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: final kernel stores into rank 0's panorama over NVLink peer memory (p2p) or NCCL send/recv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()

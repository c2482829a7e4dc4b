"""Dev tool (GPU box): the single-GPU e2e loop of bench.py (host -> host through the C ABI, `depth` async composers on `depth`
streams) with the per-stage event times of the last runs, to see where a step spends its time beyond the copy floor
(tools/e2e_copy_floor.py).

  python tools/e2e_probe.py [--steps 30] [--depth 3] [--workload cfg2]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import image_stitching_b200 as isb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--no-mask", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    isb.set_stream(stream.cuda_stream)
    rig, imgs, gains = bench.make_inputs(a.workload, 1)
    seams = bench.seam_masks_gpu(rig)
    cams = isb.cameras_from_KR(rig.Ks, rig.Rs)
    sizes = [(rig.W, rig.H)] * rig.n
    h_imgs = [torch.from_numpy(im).pin_memory().numpy() for im in imgs]
    h_gains = [torch.from_numpy(g).pin_memory().numpy() for g in gains]
    h_seams = [torch.from_numpy(s).pin_memory().numpy() for s in seams]
    comps, outs = [], []
    for _ in range(a.depth):
        c = isb.Composer(rig.warp, rig.scale, rig.nb, cache_plan=True, async_mode=True)
        _, _, roi = c.plan(cams, sizes)
        comps.append(c)
        pw, ph = roi[2], roi[3]
        outs.append((torch.zeros((ph, pw, 3), dtype=torch.uint8).pin_memory().numpy(), torch.zeros((ph, pw), dtype=torch.uint8).pin_memory().numpy()))
    streams = [torch.cuda.Stream(device=dev) for _ in range(a.depth)]
    torch.cuda.synchronize()

    def loop(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in streams:
            s.wait_event(e0)
        import time
        t0 = time.perf_counter()
        for k in range(steps):
            isb.set_stream(streams[k % a.depth].cuda_stream)
            comps[k % a.depth].run(h_imgs, h_gains, h_seams, out=outs[k % a.depth][0], out_mask=outs[k % a.depth][1])
        host_ms = (time.perf_counter() - t0) * 1e3
        for s in streams:
            stream.wait_stream(s)
        e1.record(stream)
        torch.cuda.synchronize()
        isb.set_stream(stream.cuda_stream)
        return e0.elapsed_time(e1) / steps, host_ms / steps

    loop(2 * a.depth)
    ms, host = loop(a.steps)
    print(f"depth {a.depth}: {ms:.3f} ms/step over {a.steps} steps (host submit {host:.3f} ms/step)", flush=True)
    for i, c in enumerate(comps):
        print(f"  composer {i} last run:", {k: round(v, 3) for k, v in c.timings().items()}, flush=True)


if __name__ == "__main__":
    main()

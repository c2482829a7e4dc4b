"""Dev tool (GPU box): throughput of back-to-back steps with D composers in flight on D streams (each with its own per-image
pyramids and panorama), against the single-stream chain.  Every panorama is compared with the depth-1 result.

  python tools/ab_inflight.py [--workload cfg2] [--steps 40] [--depths 1,2,3] [--priorities 0|1]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import image_stitching_b200 as isb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--div", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--depths", default="1,2,3")
    ap.add_argument("--out", default="gpurun_out/ab_inflight.json")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    main_stream = torch.cuda.current_stream()
    rig, imgs, gains = bench.make_inputs(a.workload, a.div)
    isb.set_stream(main_stream.cuda_stream)
    seams = bench.seam_masks_gpu(rig)
    cams = isb.cameras_from_KR(rig.Ks, rig.Rs)
    d_imgs = [torch.from_numpy(x).to(dev) for x in imgs]
    d_gains = [torch.from_numpy(x).to(dev) for x in gains]
    d_seams = [torch.from_numpy(x).to(dev) for x in seams]
    ref = None
    rows = []
    for depth in [int(v) for v in a.depths.split(",")]:
        streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        slots = []
        for k in range(depth):
            isb.set_stream(streams[k].cuda_stream)
            c = isb.Composer(rig.warp, rig.scale, rig.nb, async_mode=True)
            _, _, roi = c.plan(cams, [(rig.W, rig.H)] * rig.n)
            pw, ph = roi[2], roi[3]
            slots.append((c, torch.zeros((ph, pw, 3), dtype=torch.uint8, device=dev), torch.zeros((ph, pw), dtype=torch.uint8, device=dev)))
        torch.cuda.synchronize()

        def run(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_stream)
            for s in streams:
                s.wait_event(e0)
            for k in range(steps):
                c, o, m = slots[k % depth]
                isb.set_stream(streams[k % depth].cuda_stream)
                c.run(d_imgs, d_gains, d_seams, out=o, out_mask=m)
            for s in streams:
                main_stream.wait_stream(s)
            e1.record(main_stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        run(3 * depth)
        best = min(run(a.steps) for _ in range(3))
        got = [(o.cpu().numpy(), m.cpu().numpy()) for _, o, m in slots]
        if ref is None:
            ref = got[0]
        same = all(np.array_equal(g[0], ref[0]) and np.array_equal(g[1], ref[1]) for g in got)
        rows.append({"depth": depth, "ms_per_step": best, "equal_to_depth1": bool(same)})
        print(f"depth {depth}: {best:.4f} ms/step equal={same}", flush=True)
        del slots
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

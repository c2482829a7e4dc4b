"""Dev tool (GPU box): A/B of run-time switches of libisb.so (environment variables read on every run, e.g. ISB_PDL) on
one rig with the inputs built once.  Every variant's panorama is compared bit for bit with the first variant's.

  python tools/ab_env.py [--workload cfg2] [--steps 30] [--variants "ISB_PDL=0;ISB_PDL=1"]
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
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--variants", default="ISB_PDL=0;ISB_PDL=1", help="';'-separated variants, each a ' '-separated list of VAR=VALUE")
    ap.add_argument("--out", default="gpurun_out/ab_env.json")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    isb.set_stream(stream.cuda_stream)
    rig, imgs, gains = bench.make_inputs(a.workload, a.div)
    seams = bench.seam_masks_gpu(rig)
    cams = isb.cameras_from_KR(rig.Ks, rig.Rs)
    d_imgs = [torch.from_numpy(x).to(dev) for x in imgs]
    d_gains = [torch.from_numpy(x).to(dev) for x in gains]
    d_seams = [torch.from_numpy(x).to(dev) for x in seams]
    ref = None
    rows = []
    for variant in a.variants.split(";"):
        env = dict(kv.split("=", 1) for kv in variant.split())
        os.environ.update(env)
        isb.lib().isb_reload_env()
        comp = isb.Composer(rig.warp, rig.scale, rig.nb)
        _, _, roi = comp.plan(cams, [(rig.W, rig.H)] * rig.n)
        pw, ph = roi[2], roi[3]
        d_out = torch.zeros((ph, pw, 3), dtype=torch.uint8, device=dev)
        d_mask = torch.zeros((ph, pw), dtype=torch.uint8, device=dev)

        def step():
            comp.run(d_imgs, d_gains, d_seams, out=d_out, out_mask=d_mask, out_pitch=pw * 3, mask_pitch=pw)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        best = 1e9
        for _rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(a.steps):
                step()
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / a.steps)
        got = (d_out.cpu().numpy(), d_mask.cpu().numpy())
        if ref is None:
            ref = got
        same = bool(np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]))
        stages = comp.timings()
        rows.append({"variant": variant, "ms_per_step": best, "equal_to_first": same, "stages_ms": stages})
        print(f"{variant}: {best:.4f} ms/step equal={same} stages={ {k: round(v, 3) for k, v in stages.items()} }", flush=True)
        for k in env:
            os.environ.pop(k, None)
        isb.lib().isb_reload_env()
        del comp
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)
    assert all(r["equal_to_first"] for r in rows), "a variant changed the panorama"


if __name__ == "__main__":
    main()

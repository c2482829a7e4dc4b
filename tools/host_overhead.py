"""Dev tool: host-side cost of one Composer.run() call vs the device time of the step (cfg2, device-resident inputs)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import image_stitching_b200 as isb
from image_stitching_b200 import synth

div = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rig = synth.make_rig("cfg2", scale_div=div)
dev = torch.device("cuda:0")
imgs = [synth.make_image(i, rig.W, rig.H) for i in range(rig.n)]
gains = synth.make_gains(rig.n)
src = synth.seam_source_mask(rig.W, rig.H)
seams = []
for K, R in zip(rig.Ks, rig.Rs):
    Ks, ss = synth.seam_camera(K, rig.scale)
    seams.append(isb.RotationWarper(rig.warp, ss).warp(src, Ks, R, isb.INTER_NEAREST, isb.BORDER_CONSTANT)[1])
stream = torch.cuda.current_stream()
isb.set_stream(stream.cuda_stream)
comp = isb.Composer(rig.warp, rig.scale, rig.nb, cache_plan=True)
_, _, roi = comp.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
pw, ph = roi[2], roi[3]
d_imgs = [torch.from_numpy(a).to(dev) for a in imgs]
d_gains = [torch.from_numpy(g).to(dev) for g in gains]
d_seams = [torch.from_numpy(s).to(dev) for s in seams]
d_out = torch.zeros((ph, pw, 3), dtype=torch.uint8, device=dev)
d_mask = torch.zeros((ph, pw), dtype=torch.uint8, device=dev)
for _ in range(5):
    comp.run(d_imgs, d_gains, d_seams, out=d_out, out_mask=d_mask)
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N):
    comp.run(d_imgs, d_gains, d_seams, out=d_out, out_mask=d_mask)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"div={div} pano={pw}x{ph}: host enqueue {1e3*(t1-t0)/N:.3f} ms/step, total {1e3*(t2-t0)/N:.3f} ms/step")

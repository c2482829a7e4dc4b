import numpy as np, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import image_stitching_b200 as isb
from conftest import make_case, seam_masks_oracle
from oracle import oracle as orc
rig, imgs, gains, nb = make_case("cfg2", 8, 5)
seams = seam_masks_oracle(rig)
out = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
print("equal", np.array_equal(out["result16"], ref["result16"]), np.array_equal(out["mask"], ref["mask"]))

"""Generates tests/golden/*.npz from OpenCV itself (cv2 4.13.0, IPP off) - the code the reference's
compositing loop executes.  Run from the repo root in the build container:

    python tests/golden/make_golden.py

Inputs are NOT stored: they are regenerated from image_stitching_b200/synth.py (integer-only, seeded).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cv2  # noqa: E402

from conftest import make_case  # noqa: E402
from image_stitching_b200 import synth  # noqa: E402
from oracle import cv_reference as cvr  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (rig, div, nb, max_images, kind)
    "cfg2_d16_nb3": ("cfg2", 16, 3, None, "texture"),
    "cfg2_d16_nb5_checker": ("cfg2", 16, 5, None, "checker"),
    "cfg4_d8_nb5": ("cfg4", 8, 5, None, "texture"),
    "cfg3_d32_nb4": ("cfg3", 32, 4, None, "texture"),
}


def main():
    cvr.set_parity_mode(True)
    print("cv2", cv2.__version__)
    for name, (rigname, div, nb, mx, kind) in CASES.items():
        rig, imgs, gains, nb = make_case(rigname, div, nb, mx, kind)
        seams = cvr.seam_masks_cv(rig.warp, rig.scale, rig.Ks, rig.Rs, rig.W, rig.H)
        ref = cvr.compose_cv(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams, keep_stages=True)
        st = ref["stages"][0]
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            corners=np.array(ref["corners"], np.int32), sizes=np.array(ref["sizes"], np.int32),
            dst_roi=np.array(ref["dst_roi"], np.int32), result16=ref["result16"], mask=ref["mask"],
            seam_sizes=np.array([s.shape for s in seams], np.int32),
            seam_sums=np.array([int(s.astype(np.int64).sum()) for s in seams], np.int64),
            seam0=seams[0], warped0=st["img_warped"], valid0=st["valid"], mask0=st["mask"],
            cv2_version=np.array(cv2.__version__))
        print(name, ref["dst_roi"], "covered", float((ref["mask"] > 0).mean()))
    # primitive vectors: pyramids on odd shapes (weights exercise the SIMD/scalar op-order rule)
    rng = np.random.default_rng(42)
    prim = {}
    for i, (h, w) in enumerate([(37, 53), (64, 96), (5, 131)]):
        a = rng.integers(-300, 600, (h, w, 3)).astype(np.int16)
        f = (rng.random((h, w)) * (rng.random((h, w)) > 0.3)).astype(np.float32)
        prim[f"p16_{i}"] = a
        prim[f"down16_{i}"] = cv2.pyrDown(a)
        prim[f"up16_{i}"] = cv2.pyrUp(a)
        prim[f"w_{i}"] = f
        prim[f"downw_{i}"] = cv2.pyrDown(f)
    m = rng.integers(0, 256, (33, 57)).astype(np.uint8)
    prim["mask"] = m
    prim["mask_dil"] = cv2.dilate(m, None)
    prim["mask_up"] = cv2.resize(prim["mask_dil"], (453, 260), interpolation=cv2.INTER_LINEAR_EXACT)
    g = synth.make_gains(1)[0]
    prim["gain"] = g
    prim["gain_up"] = cv2.resize(g, (451, 353), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **prim)
    print("primitives written")
    simple_blenders()
    default_flow_vectors()


def default_flow_vectors():
    """cfg1-substitute: the reference's default flow (tests/test_default_flow.py) through cv2."""
    import test_default_flow as tdf
    ref = tdf.cv2_flow(cv2)
    out = ref["out"]
    arrs = dict(nb=np.array(ref["nb"]), sz=np.array(ref["sz"], np.int32), rois=np.array(ref["rois"], np.int32),
                dst_roi=np.array(out["dst_roi"], np.int32), resized0=ref["resized"][0],
                resized_sums=np.array([int(r.astype(np.int64).sum()) for r in ref["resized"]], np.int64),
                mask=out["mask"], result16=out["result16"], cv2_version=np.array(cv2.__version__))
    for i, s_ in enumerate(ref["seams"]):
        arrs[f"seam_{i}"] = s_
    np.savez_compressed(os.path.join(OUT, "default_flow.npz"), **arrs)
    print("default_flow", out["dst_roi"], "bands", ref["nb"])


def simple_blend_inputs():
    """Seeded inputs of the Blender::NO / FeatherBlender vectors (shared with the tests)."""
    rng = np.random.default_rng(77)
    corners = [(0, 0), (150, -30), (-77, 41)]
    sizes = [(300, 200), (257, 213), (190, 260)]
    imgs, masks = [], []
    for (sw, sh) in sizes:
        imgs.append(rng.integers(0, 256, (sh, sw, 3)).astype(np.int16))
        m = np.zeros((sh, sw), np.uint8)
        m[10:-10, 10:-10] = 255
        m[20:40, 20:60] = rng.integers(0, 256, (20, 40))  # grey values count as "inside" for the distance transform
        m[50:60, 50:90] = 0
        masks.append(m)
    return corners, sizes, imgs, masks


def simple_blenders():
    """Blender::NO and FeatherBlender (image_stitching.cpp:1175-1191) + createWeightMap."""
    cvr.set_parity_mode(True)
    corners, sizes, imgs, masks = simple_blend_inputs()
    roi = cv2.detail.resultRoi(corners=corners, sizes=sizes)
    out = {"roi": np.array(roi, np.int32)}
    for tag, btype, sharp in [("no", 0, 0.02), ("feather", 1, 0.02), ("feather_sharp", 1, 1 / 37.3)]:
        b = cv2.detail.Blender_createDefault(cv2.detail.Blender_NO) if btype == 0 else cv2.detail_FeatherBlender(sharp)
        b.prepare(roi)
        for img, m, c in zip(imgs, masks, corners):
            b.feed(img, m, c)
        r, rm = b.blend(None, None)
        out[tag + "_result16"], out[tag + "_mask"] = r, rm
    out["wm0"] = cv2.detail.createWeightMap(masks[0], 0.02, None)
    out["wm_full"] = cv2.detail.createWeightMap(np.full((40, 50), 255, np.uint8), 0.02, None)
    # the whole loop with blend_type feather / no on a small rig (sharpness by the reference's rule, blend_strength 5)
    rig, imgs, gains, nb = make_case("cfg2", 16, 3)
    seams = cvr.seam_masks_cv(rig.warp, rig.scale, rig.Ks, rig.Rs, rig.W, rig.H)
    for tag in ("feather", "no"):
        ref = cvr.compose_cv(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams, blend_type=tag)
        out["loop_" + tag + "_result16"], out["loop_" + tag + "_mask"] = ref["result16"], ref["mask"]
        out["loop_dst_roi"] = np.array(ref["dst_roi"], np.int32)
    out["loop_sharpness"] = np.array(cvr.feather_sharpness(ref["dst_roi"][2], ref["dst_roi"][3]), np.float32)
    np.savez_compressed(os.path.join(OUT, "simple_blend.npz"), **out)
    print("simple blenders written")


if __name__ == "__main__":
    main()

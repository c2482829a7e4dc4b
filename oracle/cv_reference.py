"""TEST INFRASTRUCTURE ONLY - the reference's CPU implementation of the compositing path.

The reference (a1q123456/image_stitching) executes this path inside OpenCV
(vcpkg `opencv4[world]`, baseline 7bc5b8cd..., not vendored under /root/reference);
its own main() cannot be built here (needs OpenCV C++ dev files + libexif).  The same
cv::detail classes are reachable through the `cv2` 4.13.0 wheel, so this module drives
them in exactly the call order of image_stitching.cpp:1086-1229:

    warper->warp(img, K, R, INTER_LINEAR, BORDER_REFLECT)          :1154
    warper->warp(mask, K, R, INTER_NEAREST, BORDER_CONSTANT)       :1157-1159
    compensator->apply(idx, corner, img_warped, mask_warped)       :1162
    img_warped.convertTo(CV_16S)                                   :1164
    dilate -> resize(INTER_LINEAR_EXACT) -> &                      :1169-1171
    blender->prepare / feed / blend                                :1173-1225
    saturate to 8U (imwrite)                                       :1228

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this file.
It is the checker and the timed CPU baseline, never part of the product path.
"""
from __future__ import annotations

import time

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def have_cv2() -> bool:
    return cv2 is not None


def set_parity_mode(on: bool = True) -> None:
    """Parity runs switch IPP off (the reference's vcpkg build has no `ipp` feature;
    only the float gain-map resize depends on it - SURVEY.md A.7)."""
    cv2.ipp.setUseIPP(not on)
    cv2.ocl.setUseOpenCL(False)


def make_warper(kind: str, scale) -> "cv2.PyRotationWarper":
    return cv2.PyRotationWarper(kind, float(scale))


def warp_rois(kind, scale, sizes_wh, Ks, Rs):
    w = make_warper(kind, scale)
    out = []
    for (sw, sh), K, R in zip(sizes_wh, Ks, Rs):
        out.append(tuple(int(v) for v in w.warpRoi((int(sw), int(sh)), K, R)))
    return out


def seam_masks_cv(kind, scale, Ks, Rs, W, H, seam_div=8):
    """Low-res seam masks the way L5 makes masks_warped[] (image_stitching.cpp:973-989),
    with a fixed source-space band standing in for the seam finder's output."""
    from image_stitching_b200 import synth
    out = []
    src = synth.seam_source_mask(W, H)
    for K, R in zip(Ks, Rs):
        Ksm, ssm = synth.seam_camera(K, scale)
        w = make_warper(kind, ssm)
        _, m = w.warp(src, Ksm, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        out.append(m)
    return out


def feather_sharpness(dst_w, dst_h, blend_strength=5.0):
    """fb->setSharpness(1.f / blend_width), blend_width = sqrt(dst_area) * blend_strength / 100 (image_stitching.cpp:1177-1190)."""
    return float(np.float32(1.0) / (np.sqrt(np.float32(dst_w * dst_h)) * np.float32(blend_strength) / np.float32(100.0)))


def warp_cv(warper, src, K, R, interp, border):
    """warper.warp(src, K, R, interp, border).  cv::remap refuses destinations of SHRT_MAX columns or more
    (imgwarp.cpp: `dst.cols < SHRT_MAX`), which a wrap-around image of a >= 32767-px-wide panorama (cfg3: 46654) hits - the
    reference itself cannot stitch such a rig.  RotationWarperBase::warp is buildMaps + remap and remap is per-pixel, so for
    those images the checker runs the same two calls with the maps cut into column chunks (identical arithmetic)."""
    h, w = src.shape[:2]
    x, y, rw, rh = warper.warpRoi((w, h), K, R)
    if rw < 32767 and rh < 32767:
        return warper.warp(src, K, R, interp, border)
    roi, xmap, ymap = warper.buildMaps((w, h), K, R)
    out = np.empty(xmap.shape + src.shape[2:], src.dtype)
    for c0 in range(0, xmap.shape[1], 16384):
        c1 = min(c0 + 16384, xmap.shape[1])
        out[:, c0:c1] = cv2.remap(src, np.ascontiguousarray(xmap[:, c0:c1]), np.ascontiguousarray(ymap[:, c0:c1]), interp,
                                  borderMode=border)
    return (int(roi[0]), int(roi[1])), out


def compose_cv(images, Ks, Rs, scale, kind, nb, gains=None, seam_masks=None, keep_stages=False,
               timings=None, blend_type="multiband", sharpness=None):
    """Mirror of the compositing loop.  Returns dict(corners, sizes, dst_roi, result16, result8, mask).
    blend_type: "multiband" (nb bands), "feather" (sharpness, default by the reference's rule) or "no"."""
    n = len(images)
    warper = make_warper(kind, scale)
    corners, sizes = [], []
    for img, K, R in zip(images, Ks, Rs):
        h, w = img.shape[:2]
        x, y, rw, rh = warper.warpRoi((w, h), K, R)
        corners.append((int(x), int(y)))
        sizes.append((int(rw), int(rh)))
    comp = None
    if gains is not None:
        comp = cv2.detail_BlocksGainCompensator(64, 64, 1)
        comp.setMatGains([np.ascontiguousarray(g, dtype=np.float32) for g in gains])
    dst_roi = cv2.detail.resultRoi(corners=corners, sizes=sizes)
    if blend_type == "multiband":
        blender = cv2.detail_MultiBandBlender(0, int(nb))
    elif blend_type == "feather":
        blender = cv2.detail_FeatherBlender(feather_sharpness(dst_roi[2], dst_roi[3]) if sharpness is None else float(sharpness))
    else:
        blender = cv2.detail.Blender_createDefault(cv2.detail.Blender_NO)
    blender.prepare(dst_roi)
    stages = []
    tm = dict(warp_img=0.0, warp_mask=0.0, gain=0.0, to16s=0.0, seam=0.0, feed=0.0, blend=0.0)
    for i, (img, K, R) in enumerate(zip(images, Ks, Rs)):
        t0 = time.perf_counter()
        corner, img_warped = warp_cv(warper, img, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        t1 = time.perf_counter()
        mask = np.full(img.shape[:2], 255, np.uint8)
        _, mask_warped = warp_cv(warper, mask, K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        t2 = time.perf_counter()
        valid = mask_warped
        if comp is not None:
            img_warped = comp.apply(i, corners[i], img_warped, mask_warped)
        t3 = time.perf_counter()
        img_warped_s = img_warped.astype(np.int16)
        t4 = time.perf_counter()
        if seam_masks is not None:
            dil = cv2.dilate(seam_masks[i], None)
            seam = cv2.resize(dil, (mask_warped.shape[1], mask_warped.shape[0]), interpolation=cv2.INTER_LINEAR_EXACT)
            mask_warped = cv2.bitwise_and(seam, mask_warped)
        t5 = time.perf_counter()
        blender.feed(img_warped_s, mask_warped, corners[i])
        t6 = time.perf_counter()
        tm["warp_img"] += t1 - t0; tm["warp_mask"] += t2 - t1; tm["gain"] += t3 - t2
        tm["to16s"] += t4 - t3; tm["seam"] += t5 - t4; tm["feed"] += t6 - t5
        if keep_stages:
            stages.append(dict(corner=tuple(int(c) for c in corner), img_warped=img_warped, valid=valid,
                               mask=mask_warped))
    t0 = time.perf_counter()
    result, result_mask = blender.blend(None, None)
    result8 = np.clip(result, 0, 255).astype(np.uint8)
    tm["blend"] = time.perf_counter() - t0
    if timings is not None:
        timings.update(tm)
    return dict(corners=corners, sizes=sizes, dst_roi=tuple(int(v) for v in dst_roi), result16=result,
                result8=result8, mask=result_mask, stages=stages)

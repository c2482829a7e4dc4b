"""cfg1-substitute (BASELINE.json configs[0]): the reference's own DEFAULT invocation of the path.

The bundled samples.zip is a git-lfs pointer, so its images cannot be stitched here.  What can be pinned is the flow the
reference runs on ANY image set when it is started without options (image_stitching.cpp:53-55, 80: work_megapix = -1,
seam_megapix = 0.1, compose_megapix = 0.4, blend_strength = 5), fed with the synthetic cfg2 cameras:

    rotate(full_img, ROTATE_180)                                            :1093-1103
    compose_scale = min(1, sqrt(compose_megapix * 1e6 / area))              :1107-1108
    warped_image_scale *= (float)compose_work_aspect; cameras *= aspect     :1115-1127
    sz = cvRound(full_size * compose_scale); warpRoi(sz, K, R)              :1130-1141
    resize(full_img, img, Size(), compose_scale, compose_scale, LINEAR_EXACT)  :1143-1146
    seam masks warped at seam scale with K * (float)seam_work_aspect        :973-989
    num_bands = ceil(log(sqrt(dst_area) * blend_strength / 100) / log 2) - 1  :1176-1183
    warp / mask / gain / 16S / seam & mask / feed / blend / saturate        :1148-1228

`-m "not gpu"`: the C oracle against OpenCV itself (cv2) step by step - the CPU-only case SURVEY.md asks for.
`-m gpu`     : the CUDA path through the C ABI against the oracle on the same inputs.
The reduced frames keep the megapixel RATIOS of a 12 MP camera (the scale factors are what the flow depends on).
"""
import math
import os

import numpy as np
import pytest

from conftest import psnr

from image_stitching_b200 import synth
from oracle import oracle as orc

COMPOSE_MEGAPIX, SEAM_MEGAPIX, BLEND_STRENGTH = 0.4, 0.1, 5.0  # image_stitching.cpp:54, 55, 80
FRAME_MP = 12.0                                                # the cfg2 camera: 4000 x 3000


def cv_round(v):
    return int(np.rint(v))  # cvRound: round half to even


def flow_scales(W, H):
    """compose_scale, seam_scale for a W x H frame standing in for a 12 MP one (work_scale = 1: work_megapix < 0)."""
    area_mp = W * H / 1e6
    k = area_mp / FRAME_MP  # the reduced rig scales the megapixel budgets with the frame
    compose_scale = min(1.0, math.sqrt(COMPOSE_MEGAPIX * k * 1e6 / (W * H)))
    seam_scale = min(1.0, math.sqrt(SEAM_MEGAPIX * k * 1e6 / (W * H)))
    return compose_scale, seam_scale


def reference_num_bands(dst_w, dst_h):
    blend_width = np.sqrt(np.float32(dst_w * dst_h)) * np.float32(BLEND_STRENGTH) / np.float32(100.0)
    return int(math.ceil(math.log(float(blend_width)) / math.log(2.0)) - 1.0)


def default_flow(rotate180, resize_fx, warp_nearest, warp_roi, compose, div=2, n=5):
    """Runs the flow of the module docstring with the given implementation of each OpenCV call; returns every
    intermediate the two sides of a parity test have to agree on."""
    rig = synth.make_rig("cfg2", scale_div=div, max_images=n)
    W, H = rig.W, rig.H
    compose_scale, seam_scale = flow_scales(W, H)
    assert abs(compose_scale - 1) > 1e-1  # the resize branch of :1143 is the one under test
    work_aspect, seam_aspect = compose_scale / 1.0, seam_scale / 1.0
    gains = synth.make_gains(rig.n)
    # seam masks at seam scale (:973-989), a fixed source-space band standing in for the seam finder's output
    seam_w, seam_h = cv_round(W * seam_scale), cv_round(H * seam_scale)
    src_mask = synth.seam_source_mask(seam_w, seam_h)
    seam_warper_scale = np.float32(float(rig.scale) * seam_aspect)  # static_cast<float>(warped_image_scale * seam_work_aspect)
    seams = []
    for K, R in zip(rig.Ks, rig.Rs):
        Ks = K.copy()
        swa = np.float32(seam_aspect)
        Ks[0, 0] *= swa; Ks[0, 2] *= swa; Ks[1, 1] *= swa; Ks[1, 2] *= swa
        seams.append(warp_nearest(rig.warp, seam_warper_scale, src_mask, Ks, R))
    # compose-scale cameras (:1115-1141): focal, ppx, ppy are doubles scaled by the double aspect, K() -> CV_32F
    scale_c = np.float32(rig.scale) * np.float32(work_aspect)  # warped_image_scale *= static_cast<float>(compose_work_aspect)
    Kc = []
    for K in rig.Ks:
        Kd = np.eye(3)
        Kd[0, 0] = float(K[0, 0]) * work_aspect
        Kd[1, 1] = float(K[1, 1]) * work_aspect
        Kd[0, 2] = float(K[0, 2]) * work_aspect
        Kd[1, 2] = float(K[1, 2]) * work_aspect
        Kc.append(Kd.astype(np.float32))
    sz = (cv_round(W * compose_scale), cv_round(H * compose_scale))
    rois = [warp_roi(rig.warp, scale_c, sz[0], sz[1], K, R) for K, R in zip(Kc, rig.Rs)]
    x0 = min(r[0] for r in rois); y0 = min(r[1] for r in rois)
    x1 = max(r[0] + r[2] for r in rois); y1 = max(r[1] + r[3] for r in rois)
    nb = reference_num_bands(x1 - x0, y1 - y0)
    # pixels: rotate + resize (:1093-1103, :1143-1146)
    rotated, resized, fulls = [], [], []
    for i in range(rig.n):
        full = synth.make_image(i, W, H)
        fulls.append(full)
        r = rotate180(full)
        rotated.append(r)
        resized.append(resize_fx(r, compose_scale))
        assert resized[-1].shape[:2] == (sz[1], sz[0])
    out = compose(resized, Kc, rig.Rs, scale_c, rig.warp, nb, gains, seams)
    return dict(rig=rig, nb=nb, rois=rois, seams=seams, rotated=rotated, resized=resized, out=out, sz=sz, scale_c=scale_c,
                fulls=fulls, Kc=Kc, gains=gains, compose_scale=compose_scale)


def oracle_flow(**kw):
    return default_flow(
        rotate180=lambda a: orc.rotate(a, 1),
        resize_fx=lambda a, f: orc.resize_linear_exact_ex(a, cv_round(a.shape[1] * f), cv_round(a.shape[0] * f), f, f),
        warp_nearest=lambda kind, s, src, K, R: orc.warp(kind, s, src, K, R, orc.NEAREST, 0)[1],
        warp_roi=lambda kind, s, w, h, K, R: tuple(orc.warp_roi(kind, s, w, h, K, R)),
        compose=orc.compose, **kw)


def assert_same_flow(a, b):
    assert a["nb"] == b["nb"] and a["sz"] == b["sz"]
    assert [tuple(r) for r in a["rois"]] == [tuple(r) for r in b["rois"]]
    for key in ("seams", "rotated", "resized"):
        for u, v in zip(a[key], b[key]):
            assert u.shape == v.shape and np.array_equal(u, v), key
    oa, ob = a["out"], b["out"]
    assert [tuple(c) for c in oa["corners"]] == [tuple(c) for c in ob["corners"]]
    assert [tuple(s) for s in oa["sizes"]] == [tuple(s) for s in ob["sizes"]]
    assert tuple(oa["dst_roi"]) == tuple(ob["dst_roi"])
    assert np.array_equal(oa["mask"], ob["mask"])
    ra = np.clip(oa["result16"], 0, 255).astype(np.uint8) if "result16" in oa else oa["result8"]
    rb = np.clip(ob["result16"], 0, 255).astype(np.uint8) if "result16" in ob else ob["result8"]
    d = np.abs(ra.astype(int) - rb.astype(int))
    assert d.max() <= 1 and psnr(ra, rb) >= 50.0  # the north-star bar ...
    return int(d.max()), bool(np.array_equal(oa.get("result16"), ob.get("result16")))


def test_reference_band_count_formula():
    # the values SURVEY.md 8(d) quotes for the full-size rigs ("which would give 8/10/7/11")
    assert reference_num_bands(20912, 2881) == 8
    assert reference_num_bands(46655, 13903) == 10
    assert reference_num_bands(10550, 2212) == 7
    assert reference_num_bands(82488, 32653) == 11


def cv2_flow(cv2, **kw):
    """The same flow through OpenCV itself (parity mode must be on: cv_reference.set_parity_mode)."""
    from oracle import cv_reference as cvr

    def warp_nearest(kind, s, src, K, R):
        return cvr.make_warper(kind, s).warp(src, K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)[1]

    def warp_roi(kind, s, w, h, K, R):
        return tuple(int(v) for v in cvr.make_warper(kind, s).warpRoi((int(w), int(h)), K, R))

    return default_flow(
        rotate180=lambda a: cv2.rotate(a, cv2.ROTATE_180),
        resize_fx=lambda a, f: cv2.resize(a, None, fx=f, fy=f, interpolation=cv2.INTER_LINEAR_EXACT),
        warp_nearest=warp_nearest, warp_roi=warp_roi, compose=cvr.compose_cv, **kw)


def test_default_flow_oracle_vs_cv2(cv2_parity):
    """CPU only: the C restatement against OpenCV on the reference's default flow (cfg1-substitute)."""
    ref = cv2_flow(cv2_parity)
    got = oracle_flow()
    dmax, exact16 = assert_same_flow(got, ref)
    # ... and the oracle's own bar: bit-exact except for the <= 1e-4 of pixels the f32 gain-map resize can move
    n_diff = int((got["out"]["result16"] != ref["out"]["result16"]).sum())
    assert n_diff <= 1e-3 * ref["out"]["result16"].size, (n_diff, dmax)
    assert 2 <= got["nb"] <= 8


def test_default_flow_oracle_vs_golden():
    """The same check without cv2: tests/golden/default_flow.npz was written by tests/golden/make_golden.py from cv2 4.13.0."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "default_flow.npz"))
    got = oracle_flow()
    assert got["nb"] == int(g["nb"]) and got["sz"] == tuple(int(v) for v in g["sz"])
    assert [tuple(r) for r in got["rois"]] == [tuple(int(v) for v in r) for r in g["rois"]]
    assert tuple(got["out"]["dst_roi"]) == tuple(int(v) for v in g["dst_roi"])
    for i, s in enumerate(got["seams"]):
        assert np.array_equal(s, g[f"seam_{i}"])
    assert np.array_equal(got["resized"][0], g["resized0"])
    assert [int(r.astype(np.int64).sum()) for r in got["resized"]] == [int(v) for v in g["resized_sums"]]
    assert np.array_equal(got["out"]["mask"], g["mask"])
    n_diff = int((got["out"]["result16"] != g["result16"]).sum())
    assert n_diff <= 1e-3 * g["result16"].size and np.abs(got["out"]["result16"].astype(int) - g["result16"].astype(int)).max() <= 1


@pytest.mark.gpu
def test_default_flow_gpu_vs_oracle():
    """The CUDA path (rotate, resize, seam-scale warp, warpRoi, fused loop) on the reference's default flow: bit-exact."""
    import image_stitching_b200 as isb

    def warp_roi(kind, s, w, h, K, R):
        return tuple(int(v) for v in isb.RotationWarper(kind, s).warpRoi((int(w), int(h)), K, R))

    got = default_flow(
        rotate180=lambda a: isb.rotate(a, isb.ROTATE_180),
        resize_fx=lambda a, f: isb.resize_linear_exact(a, fx=f, fy=f),
        warp_nearest=lambda kind, s, src, K, R: isb.RotationWarper(kind, s).warp(src, K, R, isb.INTER_NEAREST, isb.BORDER_CONSTANT)[1],
        warp_roi=warp_roi, compose=isb.compose)
    ref = oracle_flow()
    dmax, exact16 = assert_same_flow(got, ref)
    assert dmax == 0 and exact16


@pytest.mark.gpu
@pytest.mark.parametrize("blend", ["rule", "explicit"])
def test_default_flow_one_call_from_decoded_frames(blend):
    """the ingest pre-steps inside the composer.  isb_composer_run() takes the DECODED frames, applies
    rotate(ROTATE_180) and the compose-scale INTER_LINEAR_EXACT resize on the device and warps the result - the reference's
    default flow in one call, bit-exact against the oracle flow.  `rule`: the band count comes from the reference's own blender
    set-up (image_stitching.cpp:1173-1193) as well."""
    import image_stitching_b200 as isb
    ref = oracle_flow()
    rig = ref["rig"]
    kw = dict(blend_type="multiband", blend_strength=BLEND_STRENGTH) if blend == "rule" else {}
    c = isb.Composer(rig.warp, ref["scale_c"], ref["nb"] if blend == "explicit" else 0, ingest_rotate=isb.ROTATE_180,
                     compose_scale=ref["compose_scale"], **kw)
    cams = isb.cameras_from_KR(ref["Kc"], rig.Rs)
    corners, sizes, roi = c.plan(cams, [(rig.W, rig.H)] * rig.n)  # decoded sizes
    assert [tuple(r) for r in ref["rois"]] == [(cx, cy, w, h) for (cx, cy), (w, h) in zip(corners, sizes)]
    out = c.run(ref["fulls"], ref["gains"], ref["seams"], want16=True)
    assert tuple(out["dst_roi"]) == tuple(ref["out"]["dst_roi"])
    assert np.array_equal(out["mask"], ref["out"]["mask"]) and np.array_equal(out["result16"], ref["out"]["result16"])
    # portrait frames: ROTATE_90_CLOCKWISE swaps the frame's sides before the resize
    c90 = isb.Composer(rig.warp, ref["scale_c"], ref["nb"], ingest_rotate=isb.ROTATE_90_CLOCKWISE, compose_scale=ref["compose_scale"])
    tall = [np.ascontiguousarray(np.rot90(f, 1)) for f in ref["fulls"]]  # rotating these clockwise gives the landscape frames
    c90.plan(cams, [(rig.H, rig.W)] * rig.n)
    out90 = c90.run(tall, ref["gains"], ref["seams"], want16=True)
    want = orc.compose([orc.resize_linear_exact_ex(orc.rotate(t, 0), ref["sz"][0], ref["sz"][1], ref["compose_scale"], ref["compose_scale"])
                        for t in tall], ref["Kc"], rig.Rs, ref["scale_c"], rig.warp, ref["nb"], ref["gains"], ref["seams"])
    assert np.array_equal(out90["mask"], want["mask"]) and np.array_equal(out90["result16"], want["result16"])

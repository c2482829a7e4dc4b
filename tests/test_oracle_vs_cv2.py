"""CPU: every function of the C oracle against live cv2 (the library the reference's path runs in)."""
import numpy as np
import pytest
from conftest import make_case

from image_stitching_b200 import synth
from oracle import cv_reference as cvr
from oracle import oracle as orc


def _rand_cameras(n, seed, W, H):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        f = float(rng.uniform(0.4, 3.0) * W)
        K = np.array([[f, 0, W / 2 + rng.uniform(-20, 20)], [0, f * rng.uniform(0.9, 1.1), H / 2 + rng.uniform(-20, 20)],
                      [0, 0, 1]], np.float32)
        e = rng.uniform(-np.pi, np.pi, 3) * np.array([0.5, 1.0, 0.3])
        R = synth.euler_yxz_to_R(*e).astype(np.float32)
        out.append((K, R, np.float32(f * rng.uniform(0.5, 1.5))))
    return out


@pytest.mark.parametrize("kind", ["spherical", "cylindrical"])
def test_roi_and_maps(cv2_parity, kind):
    cv2 = cv2_parity
    W, H = 320, 200
    for K, R, scale in _rand_cameras(40, 11, W, H):
        w = cv2.PyRotationWarper(kind, float(scale))
        assert tuple(w.warpRoi((W, H), K, R)) == orc.warp_roi(kind, scale, W, H, K, R)
    for K, R, scale in _rand_cameras(6, 12, W, H):
        w = cv2.PyRotationWarper(kind, float(scale))
        roi = w.warpRoi((W, H), K, R)
        if roi[2] * roi[3] > 4e6:
            continue
        _, xm, ym = w.buildMaps((W, H), K, R)
        _, xo, yo = orc.build_maps(kind, scale, W, H, K, R)
        assert np.array_equal(xm.view(np.int32), xo.view(np.int32))
        assert np.array_equal(ym.view(np.int32), yo.view(np.int32))


def test_pole_roi(cv2_parity):
    cv2 = cv2_parity
    W, H = 300, 200
    K = np.array([[150, 0, 150], [0, 150, 100], [0, 0, 1]], np.float32)
    for pitch in (-1.5, -1.2, 1.2, 1.5707):
        R = synth.euler_yxz_to_R(pitch, 0.3, 0.05).astype(np.float32)
        w = cv2.PyRotationWarper("spherical", 150.0)
        assert tuple(w.warpRoi((W, H), K, R)) == orc.warp_roi("spherical", 150.0, W, H, K, R)


@pytest.mark.parametrize("kind", ["spherical", "cylindrical"])
def test_warp(cv2_parity, kind):
    cv2 = cv2_parity
    W, H = 260, 180
    img = synth.make_image(3, W, H, "checker")
    msk = np.full((H, W), 255, np.uint8)
    for K, R, scale in _rand_cameras(5, 5, W, H):
        w = cv2.PyRotationWarper(kind, float(scale))
        if np.prod(w.warpRoi((W, H), K, R)[2:]) > 3e6:
            continue
        c1, a = w.warp(img, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
        c2, b = orc.warp(kind, scale, img, K, R, orc.LINEAR, 1)
        assert tuple(c1) == c2 and np.array_equal(a, b)
        _, a = w.warp(msk, K, R, cv2.INTER_NEAREST, cv2.BORDER_CONSTANT)
        _, b = orc.warp(kind, scale, msk, K, R, orc.NEAREST, 0)
        assert np.array_equal(a, b)


def test_seam_and_gain(cv2_parity):
    cv2 = cv2_parity
    rng = np.random.default_rng(3)
    for (sw, sh, dw, dh) in [(40, 30, 320, 240), (41, 29, 327, 233), (7, 5, 50, 41), (100, 1, 333, 1), (1, 9, 5, 77)]:
        m = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
        assert np.array_equal(cv2.dilate(m, None), orc.dilate3x3(m))
        assert np.array_equal(cv2.resize(m, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT), orc.resize_linear_exact(m, dw, dh))
    img = synth.make_image(1, 300, 220)
    g = synth.make_gains(2)[1]
    comp = cv2.detail_BlocksGainCompensator(64, 64, 1)
    comp.setMatGains([g])
    a = comp.apply(0, (0, 0), img.copy(), np.full(img.shape[:2], 255, np.uint8))
    b = orc.gain_apply(img, g)
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() <= 1 and (d > 0).mean() <= 1e-4


def test_pyramids(cv2_parity):
    cv2 = cv2_parity
    rng = np.random.default_rng(9)
    for (h, w) in [(32, 32), (33, 47), (6, 200), (64, 8), (17, 3)]:
        a = rng.integers(-2000, 2000, (h, w, 3)).astype(np.int16)
        assert np.array_equal(cv2.pyrDown(a), orc.pyrdown_16s(a))
        assert np.array_equal(cv2.pyrUp(a), orc.pyrup_16s(a))
    for w in list(range(4, 60)) + [96, 131, 257]:
        f = (rng.random((9, w)) * (rng.random((9, w)) > 0.4)).astype(np.float32)
        assert np.array_equal(cv2.pyrDown(f).view(np.int32), orc.pyrdown_32f(f).view(np.int32)), w


def test_blender(cv2_parity):
    cv2 = cv2_parity
    rng = np.random.default_rng(0)
    corners = [(0, 0), (150, -30), (-77, 41)]
    sizes = [(300, 200), (257, 213), (190, 260)]
    roi = cv2.detail.resultRoi(corners=corners, sizes=sizes)
    assert tuple(roi) == orc.result_roi(corners, sizes)
    for nb in (0, 1, 3, 5, 12):
        b = cv2.detail_MultiBandBlender(0, nb)
        b.prepare(roi)
        b2 = orc.Blender(nb)
        b2.prepare(roi)
        for (cx, cy), (sw, sh) in zip(corners, sizes):
            img = rng.integers(0, 256, (sh, sw, 3)).astype(np.int16)
            m = np.zeros((sh, sw), np.uint8)
            m[10:-10, 10:-10] = 255
            m[20:40, 20:60] = rng.integers(0, 256, (20, 40))
            b.feed(img, m, (cx, cy))
            b2.feed(img, m, (cx, cy))
        r, rm = b.blend(None, None)
        r2, rm2 = b2.blend()
        assert np.array_equal(r, r2) and np.array_equal(rm, rm2), nb


@pytest.mark.parametrize("case", [("cfg2", 8, 5), ("cfg4", 8, 5), ("cfg3", 16, 4)])
def test_compose(cv2_parity, case):
    name, div, nb = case
    rig, imgs, gains, nb = make_case(name, div, nb, max_images=12)
    seams = cvr.seam_masks_cv(rig.warp, rig.scale, rig.Ks, rig.Rs, rig.W, rig.H)
    ref = cvr.compose_cv(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    out = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    assert ref["corners"] == out["corners"] and ref["sizes"] == out["sizes"] and ref["dst_roi"] == out["dst_roi"]
    assert np.array_equal(ref["mask"], out["mask"])
    assert np.array_equal(ref["result16"], out["result16"])


def test_ingest_presteps(cv2_parity):
    """rotate(90CW / 180) and resize(INTER_LINEAR_EXACT) in both the dsize and the fx/fy form (image_stitching.cpp:1093-1146)."""
    cv2 = cv2_parity
    rng = np.random.default_rng(8)
    for shape in [(37, 53, 3), (64, 40), (1, 7, 3)]:
        a = rng.integers(0, 256, shape).astype(np.uint8)
        assert np.array_equal(cv2.rotate(a, cv2.ROTATE_90_CLOCKWISE), orc.rotate(a, 0))
        assert np.array_equal(cv2.rotate(a, cv2.ROTATE_180), orc.rotate(a, 1))
    img = rng.integers(0, 256, (370, 530, 3)).astype(np.uint8)
    for fs in (0.3651483716701107, 0.5, 0.71, 0.123, 1.3):
        ref = cv2.resize(img, None, fx=fs, fy=fs, interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(ref, orc.resize_linear_exact_ex(img, ref.shape[1], ref.shape[0], fs, fs))
    for (dw, dh) in [(194, 135), (531, 371), (1000, 37), (53, 700)]:
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(ref, orc.resize_linear_exact_ex(img, dw, dh))


def test_feather_and_no_blender(cv2_parity):
    """Blender::NO and FeatherBlender incl. createWeightMap (image_stitching.cpp:1175-1191)."""
    cv2 = cv2_parity
    rng = np.random.default_rng(0)
    corners = [(0, 0), (150, -30), (-77, 41)]
    sizes = [(300, 200), (257, 213), (190, 260)]
    roi = cv2.detail.resultRoi(corners=corners, sizes=sizes)
    full = np.full((50, 60), 255, np.uint8)  # no zero pixel: the distance saturates, the weight clamps to 1
    assert np.array_equal(cv2.detail.createWeightMap(full, 0.02, None), orc.create_weight_map(full, 0.02))
    for btype, sharp in [(0, 0.02), (1, 0.02), (1, 0.1), (1, 1 / 37.3)]:
        b = cv2.detail.Blender_createDefault(cv2.detail.Blender_NO) if btype == 0 else cv2.detail_FeatherBlender(sharp)
        b.prepare(roi)
        b2 = orc.SimpleBlender(btype, sharp)
        b2.prepare(roi)
        for (cx, cy), (sw, sh) in zip(corners, sizes):
            img = rng.integers(0, 256, (sh, sw, 3)).astype(np.int16)
            m = np.zeros((sh, sw), np.uint8)
            m[10:-10, 10:-10] = 255
            m[20:40, 20:60] = rng.integers(0, 256, (20, 40))
            m[50:60, 50:90] = 0
            if btype == 1:
                assert np.array_equal(cv2.detail.createWeightMap(m, sharp, None), orc.create_weight_map(m, sharp))
            b.feed(img, m, (cx, cy))
            b2.feed(img, m, (cx, cy))
        r, rm = b.blend(None, None)
        r2, rm2 = b2.blend()
        assert np.array_equal(r, r2) and np.array_equal(rm, rm2), (btype, sharp)


def test_timelapser(cv2_parity):
    """Timelapser::createDefault(AS_IS / CROP), initialize, process, getDst (image_stitching.cpp:1194-1215)."""
    cv2 = cv2_parity
    rng = np.random.default_rng(2)
    for corners, sizes in [([(0, 0), (150, -30), (-77, 41)], [(300, 200), (257, 213), (190, 260)]),
                           ([(0, 0), (20, 10), (-15, 25)], [(300, 200), (257, 213), (290, 160)])]:
        for ttype in (0, 1):
            a = cv2.detail.Timelapser_createDefault(ttype)
            b = orc.Timelapser(ttype)
            a.initialize(corners, sizes)
            b.initialize(corners, sizes)
            for (c, (sw, sh)) in zip(corners, sizes):
                img = rng.integers(-300, 600, (sh, sw, 3)).astype(np.int16)
                a.process(img, np.ones((sh, sw), np.uint8), c)
                b.process(img, None, c)
                ref = a.getDst().get()
                assert ref.shape == b.getDst().shape and np.array_equal(ref, b.getDst()), (ttype, c)


def test_warp_backward_matches_cv2(cv2_parity):
    """RotationWarper::warpBackward (forward map per pixel with libm's atan2f / acosf, then remap)."""
    cv2 = cv2_parity
    from oracle import cv_reference as cvr
    for name, div in (("cfg2", 16), ("cfg4", 8), ("cfg3", 32)):
        rig = synth.make_rig(name, div)
        for i in (0, 1, rig.n // 2):
            img = synth.make_image(i, rig.W, rig.H)
            w = cvr.make_warper(rig.warp, rig.scale)
            K, R = rig.Ks[i], rig.Rs[i]
            _, wi = w.warp(img, K, R, cv2.INTER_LINEAR, cv2.BORDER_REFLECT)
            for interp, border in ((cv2.INTER_LINEAR, cv2.BORDER_REFLECT), (cv2.INTER_NEAREST, cv2.BORDER_CONSTANT),
                                   (cv2.INTER_LINEAR, cv2.BORDER_CONSTANT)):
                want = w.warpBackward(wi, K, R, interp, border, (rig.W, rig.H))
                got = orc.warp_backward(rig.warp, rig.scale, wi, K, R, 1 if interp == cv2.INTER_LINEAR else 0,
                                        1 if border == cv2.BORDER_REFLECT else 0, (rig.W, rig.H))
                assert np.array_equal(want, got), (name, i, interp, border)

"""GPU: the CUDA path, called through the C ABI, against the oracle (C restatement), live cv2 where present,
and the committed golden vectors.  Bars (BASELINE.json north_star): integer ROI/corner/mask geometry
bit-exact; 8-bit pixels max|d| <= 1 and PSNR >= 50 dB.  Most stages are in fact bit-exact and are tested so."""
import os

import numpy as np
import pytest
from conftest import make_case, psnr, seam_masks_oracle

import image_stitching_b200 as isb
from image_stitching_b200 import synth
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAX_ABS = 1       # north_star tolerance on 8-bit output
MIN_PSNR = 50.0   # dB


def _cams(seed, n, W, H):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        f = float(rng.uniform(0.5, 2.5) * W)
        K = np.array([[f, 0, W / 2 + rng.uniform(-9, 9)], [0, f * rng.uniform(0.95, 1.05), H / 2 + rng.uniform(-9, 9)],
                      [0, 0, 1]], np.float32)
        e = rng.uniform(-np.pi, np.pi, 3) * np.array([0.45, 1.0, 0.2])
        out.append((K, synth.euler_yxz_to_R(*e).astype(np.float32), np.float32(f * rng.uniform(0.6, 1.4))))
    return out


def test_device_present_and_native_library_loaded():
    assert isb.device_count() > 0, "GPU tests need a CUDA device: there is no CPU fallback"
    assert b"sm_100a" in isb.lib().isb_version()


@pytest.mark.parametrize("kind", ["spherical", "cylindrical"])
def test_build_maps_bit_exact(kind):
    W, H = 300, 200
    for K, R, scale in _cams(21, 6, W, H):
        w = isb.RotationWarper(kind, scale)
        roi = w.warpRoi((W, H), K, R)
        if roi[2] * roi[3] > 6e6:
            continue
        r2, xm, ym = w.buildMaps((W, H), K, R)
        r3, xo, yo = orc.build_maps(kind, scale, W, H, K, R)
        assert r2 == r3 == roi
        assert np.array_equal(xm.view(np.int32), xo.view(np.int32))
        assert np.array_equal(ym.view(np.int32), yo.view(np.int32))


@pytest.mark.parametrize("kind", ["spherical", "cylindrical"])
@pytest.mark.parametrize("content", ["texture", "checker"])
def test_warp_bit_exact(kind, content):
    W, H = 280, 190
    img = synth.make_image(2, W, H, content)
    msk = np.full((H, W), 255, np.uint8)
    for K, R, scale in _cams(22, 5, W, H):
        w = isb.RotationWarper(kind, scale)
        if np.prod(w.warpRoi((W, H), K, R)[2:]) > 5e6:
            continue
        c1, a = w.warp(img, K, R, isb.INTER_LINEAR, isb.BORDER_REFLECT)
        c2, b = orc.warp(kind, scale, img, K, R, orc.LINEAR, 1)
        assert c1 == c2 and np.array_equal(a, b)
        c1, a = w.warp(msk, K, R, isb.INTER_NEAREST, isb.BORDER_CONSTANT)
        c2, b = orc.warp(kind, scale, msk, K, R, orc.NEAREST, 0)
        assert c1 == c2 and np.array_equal(a, b)
        # the cross combinations the warper interface also allows
        _, a = w.warp(img, K, R, isb.INTER_LINEAR, isb.BORDER_CONSTANT)
        _, b = orc.warp(kind, scale, img, K, R, orc.LINEAR, 0)
        assert np.array_equal(a, b)
        _, a = w.warp(img[:, :, 0].copy(), K, R, isb.INTER_NEAREST, isb.BORDER_REFLECT)
        _, b = orc.warp(kind, scale, img[:, :, 0].copy(), K, R, orc.NEAREST, 1)
        assert np.array_equal(a, b)


def test_wrap_around_and_pole_images():
    rig = synth.make_rig("cfg2", 16)
    img = synth.make_image(0, rig.W, rig.H)
    w = isb.RotationWarper("spherical", rig.scale)
    c1, a = w.warp(img, rig.Ks[0], rig.Rs[0], isb.INTER_LINEAR, isb.BORDER_REFLECT)  # straddles u = +-pi*scale
    c2, b = orc.warp("spherical", rig.scale, img, rig.Ks[0], rig.Rs[0], orc.LINEAR, 1)
    assert a.shape[1] > 4 * rig.W and c1 == c2 and np.array_equal(a, b)
    R = synth.euler_yxz_to_R(-1.45, 0.4, 0.02).astype(np.float32)  # looks at a pole
    c1, a = w.warp(img, rig.Ks[0], R, isb.INTER_LINEAR, isb.BORDER_REFLECT)
    c2, b = orc.warp("spherical", rig.scale, img, rig.Ks[0], R, orc.LINEAR, 1)
    assert c1 == c2 and np.array_equal(a, b)


def test_gain_apply_and_seam_mask():
    img = synth.make_image(1, 451, 353)
    gains = synth.make_gains(3)
    comp = isb.BlocksGainCompensator(64, 64, 1)
    comp.setMatGains(gains)
    assert np.array_equal(comp.getMatGain(2), gains[2])
    for i in range(3):
        assert np.array_equal(comp.apply(i, (0, 0), img), orc.gain_apply(img, gains[i]))
    with pytest.raises(isb.IsbError):
        comp.apply(7, (0, 0), img)
    p = np.load(os.path.join(GOLD, "primitives.npz"))
    valid = np.full((260, 453), 255, np.uint8)
    valid[:, :40] = 0
    assert np.array_equal(isb.seam_mask_apply(p["mask"], valid), p["mask_up"] & valid)
    rng = np.random.default_rng(4)
    for (sw, sh, dw, dh) in [(41, 29, 327, 233), (7, 5, 50, 41), (100, 1, 333, 1), (1, 9, 5, 77)]:
        m = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
        v = rng.integers(0, 2, (dh, dw)).astype(np.uint8) * 255
        assert np.array_equal(isb.seam_mask_apply(m, v), orc.resize_linear_exact(orc.dilate3x3(m), dw, dh) & v)


@pytest.mark.parametrize("nb", [0, 1, 3, 5, 12])
def test_blender_feed_blend_bit_exact(nb):
    rng = np.random.default_rng(nb)
    corners = [(0, 0), (150, -30), (-77, 41), (-60, -20)]
    sizes = [(300, 200), (257, 213), (190, 260), (31, 17)]
    a, b = isb.MultiBandBlender(0, nb), orc.Blender(nb)
    a.prepare(corners, sizes)
    b.prepare(orc.result_roi(corners, sizes))
    for (cx, cy), (sw, sh) in zip(corners, sizes):
        img = rng.integers(-50, 300, (sh, sw, 3)).astype(np.int16)
        m = np.zeros((sh, sw), np.uint8)
        m[3:-3, 3:-3] = 255
        m[5:12, 4:20] = rng.integers(0, 256, (7, 16))  # grey seam-edge weights
        a.feed(img, m, (cx, cy))
        b.feed(img, m, (cx, cy))
    r1, m1 = a.blend()
    r2, m2 = b.blend()
    assert np.array_equal(m1, m2)
    assert np.array_equal(r1, r2)


def test_blender_contract_errors():
    b = isb.MultiBandBlender(0, 3)
    with pytest.raises(isb.IsbError):  # feed before prepare
        b.feed(np.zeros((8, 8, 3), np.int16), np.zeros((8, 8), np.uint8), (0, 0))
    b.prepare((0, 0, 64, 64))
    with pytest.raises(isb.IsbError):  # img must be CV_16SC3 (blenders.cpp:365)
        b.feed(np.zeros((8, 8, 3), np.float32), np.zeros((8, 8), np.uint8), (0, 0))
    with pytest.raises(isb.IsbError):  # mask must be CV_8U (blenders.cpp:366)
        b.feed(np.zeros((8, 8, 3), np.int16), np.zeros((8, 8), np.float32), (0, 0))
    # an image-less blend gives an all-zero panorama and mask
    r, m = b.blend()
    assert r.shape == (64, 64, 3) and not r.any() and not m.any()
    with pytest.raises(isb.IsbError):  # single use per prepare()
        b.blend()
    w = isb.RotationWarper("spherical", 10.0)
    with pytest.raises(isb.IsbError):
        isb.RotationWarper("fisheye", 10.0)
    with pytest.raises(isb.IsbError):
        w.warpRoi((0, 10), np.eye(3), np.eye(3))


def _check(out, ref, exact16=True):
    assert out["corners"] == ref["corners"] and out["sizes"] == ref["sizes"] and tuple(out["dst_roi"]) == tuple(ref["dst_roi"])
    assert np.array_equal(out["mask"], ref["mask"])  # integer mask geometry: bit-exact
    r8 = np.clip(ref["result16"], 0, 255).astype(np.uint8)
    d = np.abs(out["result8"].astype(int) - r8.astype(int))
    assert d.max() <= MAX_ABS, int(d.max())
    assert psnr(out["result8"], r8) >= MIN_PSNR
    if exact16:
        assert np.array_equal(out["result16"], ref["result16"])


@pytest.mark.parametrize("case", [("cfg2", 8, 5, "texture"), ("cfg2", 16, 3, "checker"), ("cfg4", 8, 5, "texture"),
                                  ("cfg3", 16, 4, "texture"), ("cfg5", 16, 4, "texture")])
def test_compose_vs_oracle(case):
    name, div, nb, kind = case
    rig, imgs, gains, nb = make_case(name, div, nb, max_images=14, kind=kind)
    seams = seam_masks_oracle(rig)
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    out = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    _check(out, ref)
    # no gains / no seam masks (both optional in the loop)
    ref = orc.compose(imgs[:3], rig.Ks[:3], rig.Rs[:3], rig.scale, rig.warp, nb)
    out = isb.compose(imgs[:3], rig.Ks[:3], rig.Rs[:3], rig.scale, rig.warp, nb)
    _check(out, ref)


def test_compose_large_tiles():
    """Tiles big enough for the TMA-staged pyrDown boxes (136 x 67) and several CTAs per tile, incl. the wrap image."""
    rig, imgs, gains, nb = make_case("cfg2", 4, 5)
    seams = seam_masks_oracle(rig)
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    out = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    _check(out, ref)
    rig, imgs, gains, nb = make_case("cfg4", 4, 5)
    seams = seam_masks_oracle(rig)
    _check(isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams),
           orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams))


@pytest.mark.parametrize("name", ["cfg2_d16_nb3", "cfg2_d16_nb5_checker", "cfg4_d8_nb5", "cfg3_d32_nb4"])
def test_compose_vs_golden(name):
    cases = {"cfg2_d16_nb3": ("cfg2", 16, 3, "texture"), "cfg2_d16_nb5_checker": ("cfg2", 16, 5, "checker"),
             "cfg4_d8_nb5": ("cfg4", 8, 5, "texture"), "cfg3_d32_nb4": ("cfg3", 32, 4, "texture")}
    rigname, div, nb, kind = cases[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    rig, imgs, gains, nb = make_case(rigname, div, nb, kind=kind)
    # seam masks through OUR warper at seam scale - must reproduce cv2's
    src = synth.seam_source_mask(rig.W, rig.H)
    seams = []
    for K, R in zip(rig.Ks, rig.Rs):
        Ks, ss = synth.seam_camera(K, rig.scale)
        seams.append(isb.RotationWarper(rig.warp, ss).warp(src, Ks, R, isb.INTER_NEAREST, isb.BORDER_CONSTANT)[1])
    assert [s.shape for s in seams] == [tuple(v) for v in g["seam_sizes"]]
    assert np.array_equal(seams[0], g["seam0"])
    out = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    ref = dict(corners=[tuple(v) for v in g["corners"]], sizes=[tuple(v) for v in g["sizes"]], dst_roi=tuple(g["dst_roi"]),
               mask=g["mask"], result16=g["result16"])
    _check(out, ref)


def test_compose_with_device_pointers_and_plan_cache():
    torch = pytest.importorskip("torch")
    rig, imgs, gains, nb = make_case("cfg2", 16, 4)
    seams = seam_masks_oracle(rig)
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    c = isb.Composer(rig.warp, rig.scale, nb, cache_plan=True)
    c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
    dimgs = [torch.from_numpy(im).cuda() for im in imgs]
    x, y, w, h = c.dst_roi
    o8 = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    om = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    o16 = torch.zeros((h, w, 3), dtype=torch.int16, device="cuda")
    for _ in range(2):  # second run reuses the cached plan and all buffers
        c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
        c.run(dimgs, gains, seams, out=o8, out_mask=om, out16=o16)
        torch.cuda.synchronize()
        _check(dict(corners=c.corners, sizes=c.sizes, dst_roi=c.dst_roi, mask=om.cpu().numpy(), result8=o8.cpu().numpy(),
                    result16=o16.cpu().numpy()), ref)
    t = c.timings()
    assert set(t) == {"h2d", "warp", "pyrdown", "blend", "d2h"} and t["warp"] > 0
    bm = c.byte_model()
    valid = sum(int((orc.warp(rig.warp, rig.scale, np.full((rig.H, rig.W), 255, np.uint8), K, R, orc.NEAREST, 0)[1] > 0).sum())
                for K, R in zip(rig.Ks, rig.Rs))
    assert bm["M"] == valid and bm["S"] == rig.n * rig.W * rig.H and bm["Ap"] == w * h


@pytest.mark.parametrize("strips", [2, 3, 8])
def test_strip_sharding_is_bit_exact(strips):
    """N logical strips on one GPU reproduce the unsharded panorama bit for bit."""
    rig, imgs, gains, nb = make_case("cfg3", 16, 3)
    seams = seam_masks_oracle(rig)
    full = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    h, w = full["mask"].shape
    out8, outm, out16 = np.zeros((h, w, 3), np.uint8), np.zeros((h, w), np.uint8), np.zeros((h, w, 3), np.int16)
    rows = []
    for i in range(strips):
        c = isb.Composer(rig.warp, rig.scale, nb, strip_index=i, strip_count=strips)
        c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
        r = c.run(imgs, gains, seams, out=out8, out_mask=outm, out16=out16)
        rows.append(r["strip_rows"])
    assert rows[0][0] == 0 and rows[-1][1] == h and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    assert np.array_equal(outm, full["mask"]) and np.array_equal(out16, full["result16"]) and np.array_equal(out8, full["result8"])


def test_edge_cases_tiny_single_and_empty_masks():
    """Edge cases the domain has: a single image, tiny sources (vectorised sampler disabled), an all-zero seam mask
    (image contributes nothing), a missing gain / seam entry, zero bands, odd panorama widths."""
    rng = np.random.default_rng(11)
    # (a) one tiny image, nb = 0 and nb = 2
    K = np.array([[9, 0, 2.5], [0, 9, 2], [0, 0, 1]], np.float32)
    R = synth.euler_yxz_to_R(0.05, 0.3, -0.02).astype(np.float32)
    img = rng.integers(0, 256, (4, 5, 3)).astype(np.uint8)
    for nb in (0, 2):
        _check(isb.compose([img], [K], [R], 9.0, "spherical", nb), orc.compose([img], [K], [R], 9.0, "spherical", nb))
    # (b) three images: one fully masked out by its seam mask, one without gain, one without seam mask
    rig, imgs, gains, nb = make_case("cfg4", 16, 3, max_images=3)
    seams = seam_masks_oracle(rig)
    seams[1] = np.zeros_like(seams[1])
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    out = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    _check(out, ref)
    assert (out["mask"] == 0).any()  # the masked-out image leaves a hole
    c = isb.Composer(rig.warp, rig.scale, nb)
    c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
    out = c.run(imgs, [gains[0], None, gains[2]], [seams[0], seams[1], None], want16=True)
    # the oracle takes per-image optional inputs as identity gain / all-255 seam
    g1 = np.ones((1, 1), np.float32)
    s2 = np.full((4, 4), 255, np.uint8)
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, [gains[0], g1, gains[2]], [seams[0], seams[1], s2])
    _check(out, ref)
    # (c) cylindrical with an odd-width panorama buffer
    rig, imgs, gains, nb = make_case("cfg4", 16, 2, max_images=2)
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb)
    out = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb)
    _check(out, ref)


def test_cfg4_cameras_through_serializer(tmp_path):
    """BASELINE config 4: fixed cameras written to / loaded from cams.data (lossy 6-digit text, serializer.cpp:113-167);
    the composer must agree with the oracle on the LOADED cameras, frame after frame, with the cached plan."""
    rig, imgs, gains, nb = make_case("cfg4", 8, 5)
    path = str(tmp_path / "cams.data")
    isb.serializeCameraParams(isb.cameras_from_KR(rig.Ks, rig.Rs), path)
    cams = isb.deserializeCameraParams(path)
    Ks = [c.K() for c in cams]
    Rs = [np.array(list(c.R), np.float32).reshape(3, 3) for c in cams]
    assert not all(np.array_equal(a, b) for a, b in zip(Rs, rig.Rs))  # the text format really is lossy
    seams = []
    src = synth.seam_source_mask(rig.W, rig.H)
    for K, R in zip(Ks, Rs):
        Ksm, ss = synth.seam_camera(K, rig.scale)
        seams.append(orc.warp(rig.warp, ss, src, Ksm, R, orc.NEAREST, 0)[1])
    c = isb.Composer(rig.warp, rig.scale, nb, cache_plan=True)
    for frame in range(3):  # new pixels each frame, geometry / masks / gains fixed
        frames = [synth.make_image(100 * frame + i, rig.W, rig.H) for i in range(rig.n)]
        c.plan(cams, [(rig.W, rig.H)] * rig.n)
        out = c.run(frames, gains, seams, want16=True)
        _check(out, orc.compose(frames, Ks, Rs, rig.scale, rig.warp, nb, gains, seams))


def test_ingest_presteps_rotate_and_resize():
    """rotate(90CW/180) + resize(compose_scale, INTER_LINEAR_EXACT), bit-exact."""
    rng = np.random.default_rng(8)
    for shape in [(37, 53, 3), (64, 40), (1, 7, 3), (300, 201, 3)]:
        a = rng.integers(0, 256, shape).astype(np.uint8)
        assert np.array_equal(isb.rotate(a, isb.ROTATE_90_CLOCKWISE), orc.rotate(a, 0))
        assert np.array_equal(isb.rotate(a, isb.ROTATE_180), orc.rotate(a, 1))
    img = synth.make_image(5, 530, 370)
    for fs in (0.3651483716701107, 0.5, 0.71, 0.123, 1.3):
        out = isb.resize_linear_exact(img, fx=fs, fy=fs)
        assert np.array_equal(out, orc.resize_linear_exact_ex(img, out.shape[1], out.shape[0], fs, fs))
    for (dw, dh) in [(194, 135), (531, 371), (1000, 37), (53, 700)]:
        assert np.array_equal(isb.resize_linear_exact(img, (dw, dh)), orc.resize_linear_exact_ex(img, dw, dh))
    m = rng.integers(0, 256, (33, 57)).astype(np.uint8)
    assert np.array_equal(isb.resize_linear_exact(m, (453, 260)), orc.resize_linear_exact(m, 453, 260))


def test_seam_scale_auxiliary_warp():
    """SURVEY.md 8(f) rank 1 (image_stitching.cpp:973-995): the low-resolution warp of every image and mask that feeds
    the (CPU) exposure compensator and seam finder - same warper, K scaled by seam_work_aspect, scale * aspect."""
    rig, imgs, _, _ = make_case("cfg2", 4, 5, max_images=3)
    aspect = 1.0 / synth.SEAM_DIV
    for img, K, R in zip(imgs, rig.Ks, rig.Rs):
        small = isb.resize_linear_exact(img, fx=aspect, fy=aspect)   # seam_scale resize (:604-622)
        assert np.array_equal(small, orc.resize_linear_exact_ex(img, small.shape[1], small.shape[0], aspect, aspect))
        Ks, ss = synth.seam_camera(K, rig.scale)
        w = isb.RotationWarper(rig.warp, ss)
        c1, a = w.warp(small, Ks, R, isb.INTER_LINEAR, isb.BORDER_REFLECT)
        c2, b = orc.warp(rig.warp, ss, small, Ks, R, orc.LINEAR, 1)
        assert c1 == c2 and np.array_equal(a, b)
        m = np.full(small.shape[:2], 255, np.uint8)
        c1, a = w.warp(m, Ks, R, isb.INTER_NEAREST, isb.BORDER_CONSTANT)
        c2, b = orc.warp(rig.warp, ss, m, Ks, R, orc.NEAREST, 0)
        assert c1 == c2 and np.array_equal(a, b)


# ---- Blender::NO and FeatherBlender (image_stitching.cpp:1175-1191) ---------------------
def test_create_weight_map_bit_exact():
    rng = np.random.default_rng(21)
    g = np.load(os.path.join(GOLD, "simple_blend.npz"))
    from test_oracle_golden import simple_blend_inputs
    _, _, _, masks = simple_blend_inputs()
    assert np.array_equal(isb.createWeightMap(masks[0], 0.02), g["wm0"])
    assert np.array_equal(isb.createWeightMap(np.full((40, 50), 255, np.uint8), 0.02), g["wm_full"])  # no zero pixel
    for (h, w) in [(1, 1), (1, 77), (91, 1), (33, 47), (200, 333), (64, 1025)]:
        m = (rng.random((h, w)) > 0.02).astype(np.uint8) * rng.integers(1, 256, (h, w)).astype(np.uint8)
        for sharp in (0.02, 0.5, 1 / 390.0):
            assert np.array_equal(isb.createWeightMap(m, sharp), orc.create_weight_map(m, sharp)), (h, w, sharp)
    with pytest.raises(isb.IsbError):  # the saturated "no zero pixel" distance must still clamp to 1
        isb.createWeightMap(np.full((4, 4), 255, np.uint8), 1e-6)


@pytest.mark.parametrize("tag,btype,sharp", [("no", 0, 0.02), ("feather", 1, 0.02), ("feather_sharp", 1, 1 / 37.3)])
def test_simple_blenders_bit_exact(tag, btype, sharp):
    from test_oracle_golden import simple_blend_inputs
    g = np.load(os.path.join(GOLD, "simple_blend.npz"))
    corners, sizes, imgs, masks = simple_blend_inputs()
    b = isb.Blender_createDefault(btype)
    if btype == isb.BLENDER_FEATHER:
        assert abs(b.sharpness() - 0.02) < 1e-9  # FeatherBlender(0.02f) default
        b.setSharpness(sharp)
    b.prepare(corners, sizes)
    for img, m, c in zip(imgs, masks, corners):
        b.feed(img, m, c)
    r, rm = b.blend()
    assert np.array_equal(rm, g[tag + "_mask"])
    assert np.array_equal(r, g[tag + "_result16"])
    with pytest.raises(isb.IsbError):  # single use per prepare()
        b.blend()


def test_simple_blenders_vs_oracle_wraparound_and_contract():
    """int16 wrap-around of the feather accumulator, negative inputs, device-resident inputs, ROI assertion."""
    rng = np.random.default_rng(5)
    roi = (-20, -10, 400, 300)
    for btype in (isb.BLENDER_NO, isb.BLENDER_FEATHER):
        a = isb.Blender_createDefault(btype)
        o = orc.SimpleBlender(btype, 0.02)
        a.prepare(roi)
        o.prepare(roi)
        for k in range(6):  # six overlapping full-weight images of large values wrap the int16 sum
            w, h = 300, 250
            img = rng.integers(-30000, 30000, (h, w, 3)).astype(np.int16)
            m = np.full((h, w), 255, np.uint8)
            m[:, :3] = 0
            tl = (-20 + 7 * k, -10 + 5 * k)
            a.feed(img, m, tl)
            o.feed(img, m, tl)
        r1, m1 = a.blend()
        r2, m2 = o.blend()
        assert np.array_equal(m1, m2) and np.array_equal(r1, r2), btype
    b = isb.FeatherBlender()
    with pytest.raises(isb.IsbError):  # feed before prepare
        b.feed(np.zeros((8, 8, 3), np.int16), np.zeros((8, 8), np.uint8), (0, 0))
    b.prepare((0, 0, 32, 32))
    with pytest.raises(isb.IsbError):  # image outside dst_roi_
        b.feed(np.zeros((8, 8, 3), np.int16), np.zeros((8, 8), np.uint8), (30, 0))
    torch = pytest.importorskip("torch")
    img = rng.integers(0, 256, (16, 16, 3)).astype(np.int16)
    m = np.full((16, 16), 255, np.uint8)
    m[0, :] = 0
    b.feed(torch.from_numpy(img).cuda(), torch.from_numpy(m).cuda(), (4, 4))  # device pointers are used in place
    o = orc.SimpleBlender(1, 0.02)
    o.prepare((0, 0, 32, 32))
    o.feed(img, m, (4, 4))
    r1, m1 = b.blend()
    r2, m2 = o.blend()
    assert np.array_equal(m1, m2) and np.array_equal(r1, r2)


@pytest.mark.parametrize("pad,staged", [(0, True), (2, False), (1, True), (1, False), (6, True), (13, True)])
def test_compose_8bit_device_output_paths(pad, staged, monkeypatch):
    """8UC3 + mask only, device-resident, rows wider than the panorama by `pad` bytes beyond a multiple of 16: the packed
    2-byte stores (even pitches), the byte stores (odd pitches) and - forced here on local memory, normally reserved for the
    strip-sharded peer-memory output - the shared-memory staged 16-byte stores at every alignment, against the generic path
    that the 16SC3 request of the other tests takes."""
    torch = pytest.importorskip("torch")
    if staged:
        monkeypatch.setenv("ISB_STAGED_STORES", "1")
    isb.lib().isb_reload_env()  # the switches are read once per process
    rig, imgs, gains, nb = make_case("cfg2", 8, 5)
    seams = seam_masks_oracle(rig)
    ref = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)  # generic stores (16SC3 requested)
    c = isb.Composer(rig.warp, rig.scale, nb)
    c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
    x, y, w, h = c.dst_roi
    p8 = (w * 3 + 15) // 16 * 16 + pad
    pm = (w + 15) // 16 * 16 + pad
    o8 = torch.full((h, p8), 7, dtype=torch.uint8, device="cuda")
    om = torch.full((h, pm), 7, dtype=torch.uint8, device="cuda")
    c.run([torch.from_numpy(im).cuda() for im in imgs], gains, seams, out=o8, out_mask=om, out_pitch=p8, mask_pitch=pm)
    torch.cuda.synchronize()
    o8, om = o8.cpu().numpy(), om.cpu().numpy()
    monkeypatch.undo()
    isb.lib().isb_reload_env()  # back to the production switches for the tests that follow
    assert np.array_equal(o8[:, :w * 3].reshape(h, w, 3), ref["result8"]) and np.array_equal(om[:, :w], ref["mask"])
    assert (o8[:, w * 3:] == 7).all() and (om[:, w:] == 7).all()  # nothing written beyond the panorama's columns


@pytest.mark.parametrize("base_off,pad", [(1, 0), (1, 1), (0, 3), (3, 5)])
def test_compose_8bit_odd_addresses_fast_path(base_off, pad):
    """A tightly packed panorama of odd width (cfg3: 46655 columns) has rows at odd addresses.  The pipelined level-0 blend
    serves it too (rows at odd addresses split the first and last colour byte off their 16-bit stores): odd base pointer with
    an even pitch (every row odd), odd base with an odd pitch, even base with an odd pitch (rows alternate)."""
    torch = pytest.importorskip("torch")
    rig, imgs, gains, nb = make_case("cfg2", 8, 5)
    seams = seam_masks_oracle(rig)
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    c = isb.Composer(rig.warp, rig.scale, nb)
    c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
    x, y, w, h = c.dst_roi
    p8, pm = w * 3 + pad, w + pad
    b8 = torch.full((h * p8 + 64,), 7, dtype=torch.uint8, device="cuda")
    bm = torch.full((h * pm + 64,), 7, dtype=torch.uint8, device="cuda")
    o8, om = b8[base_off:], bm[base_off:]
    assert o8.data_ptr() % 2 == base_off % 2
    c.run([torch.from_numpy(im).cuda() for im in imgs], gains, seams, out=o8, out_mask=om, out_pitch=p8, mask_pitch=pm)
    torch.cuda.synchronize()
    b8, bm = b8.cpu().numpy(), bm.cpu().numpy()
    g8 = b8[base_off:base_off + h * p8].reshape(h, p8)
    gm = bm[base_off:base_off + h * pm].reshape(h, pm)
    assert np.array_equal(g8[:, :w * 3].reshape(h, w, 3), ref["result8"]) and np.array_equal(gm[:, :w], ref["mask"])
    assert (g8[:, w * 3:] == 7).all() and (gm[:, w:] == 7).all() and (b8[:base_off] == 7).all() and (bm[:base_off] == 7).all()
    assert (b8[base_off + h * p8:] == 7).all() and (bm[base_off + h * pm:] == 7).all()  # nothing written outside the rows


@pytest.mark.parametrize("tag", ["feather", "no"])
def test_loop_with_simple_blenders(tag):
    """The compositing loop call by call through the C ABI with blend_type feather / no (image_stitching.cpp:1175-1191):
    bit-exact against the same loop composed from the oracle, and within the north-star bar of the cv2 loop vectors."""
    from test_oracle_golden import oracle_loop_with_blender
    g = np.load(os.path.join(GOLD, "simple_blend.npz"))
    rig, imgs, gains, nb = make_case("cfg2", 16, 3)
    seams = seam_masks_oracle(rig)
    sharp = float(g["loop_sharpness"])
    gpu_b = isb.FeatherBlender(sharp) if tag == "feather" else isb.Blender_createDefault(isb.BLENDER_NO)
    out = isb.compose_with_blender(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, gpu_b, gains, seams)
    ref = oracle_loop_with_blender(rig, imgs, gains, seams, orc.SimpleBlender(1 if tag == "feather" else 0, sharp))
    assert tuple(out["dst_roi"]) == tuple(ref["dst_roi"]) == tuple(g["loop_dst_roi"])
    assert np.array_equal(out["mask"], ref["mask"]) and np.array_equal(out["result16"], ref["result16"])
    assert np.array_equal(out["mask"], g["loop_" + tag + "_mask"])
    cv8 = np.clip(g["loop_" + tag + "_result16"], 0, 255).astype(np.uint8)
    assert np.abs(out["result8"].astype(int) - cv8.astype(int)).max() <= MAX_ABS and psnr(out["result8"], cv8) >= MIN_PSNR


def test_warp_backward_bit_exact():
    """isb_warper_warp_backward (RotationWarper::warpBackward): the forward map's atan2f / acosf run per pixel on
    the device with glibc's float algorithms (csrc/glibc_math.cuh), so the maps - and with them the resampled frame - are the
    oracle's (which is pinned against cv2's warpBackward in tests/test_oracle_vs_cv2.py) bit for bit."""
    for name, div in (("cfg2", 8), ("cfg4", 4), ("cfg3", 16)):
        rig = synth.make_rig(name, div)
        for i in (0, 1, rig.n // 2, rig.n - 1):
            img = synth.make_image(i, rig.W, rig.H)
            K, R = rig.Ks[i], rig.Rs[i]
            w = isb.RotationWarper(rig.warp, rig.scale)
            _, wi = w.warp(img, K, R, isb.INTER_LINEAR, isb.BORDER_REFLECT)
            _, wm = w.warp(np.full(img.shape[:2], 255, np.uint8), K, R, isb.INTER_NEAREST, isb.BORDER_CONSTANT)
            for interp, border in ((isb.INTER_LINEAR, isb.BORDER_REFLECT), (isb.INTER_NEAREST, isb.BORDER_CONSTANT),
                                   (isb.INTER_LINEAR, isb.BORDER_CONSTANT), (isb.INTER_NEAREST, isb.BORDER_REFLECT)):
                want = orc.warp_backward(rig.warp, rig.scale, wi, K, R, 1 if interp == isb.INTER_LINEAR else 0,
                                         1 if border == isb.BORDER_REFLECT else 0, (rig.W, rig.H))
                got = w.warpBackward(wi, K, R, interp, border, (rig.W, rig.H))
                assert np.array_equal(want, got), (name, i, interp, border)
            # the mask (8UC1) comes back as the full frame wherever the round trip stays inside the warped ROI
            back = w.warpBackward(wm, K, R, isb.INTER_NEAREST, isb.BORDER_CONSTANT, (rig.W, rig.H))
            assert np.array_equal(back, orc.warp_backward(rig.warp, rig.scale, wm, K, R, 0, 0, (rig.W, rig.H)))
            assert (back == 255).mean() > 0.98
    with pytest.raises(isb.IsbError):  # CV_Assert on the source size
        isb.RotationWarper("spherical", 100.0).warpBackward(np.zeros((10, 10), np.uint8), np.eye(3), np.eye(3), 0, 0, (64, 48))


@pytest.mark.parametrize("tag", ["feather", "no", "multiband"])
def test_fused_composer_serves_all_three_blenders(tag):
    """isb_config.use_blend_rule: the C composer applies the reference's own blender set-up (image_stitching.cpp:1173-1193:
    blend_width from the panorama area and blend_strength -> Blender::NO / FeatherBlender sharpness / band count) and runs the
    whole loop on the device for every blender type - bit-exact against the loop composed from the oracle."""
    from test_oracle_golden import oracle_loop_with_blender
    rig, imgs, gains, nb = make_case("cfg2", 16, 3)
    seams = seam_masks_oracle(rig)
    c = isb.Composer(rig.warp, rig.scale, 99, blend_type=tag, blend_strength=5.0)  # num_bands is ignored under the rule
    _, _, roi = c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
    out = c.run(imgs, gains, seams, want16=True)
    bw = float(np.sqrt(np.float32(roi[2] * roi[3])) * np.float32(5.0) / np.float32(100.0))
    assert bw >= 1
    if tag == "multiband":
        bands = int(np.ceil(np.log(bw) / np.log(2.0)) - 1.0)
        assert bands == isb.num_bands_for(roi[2], roi[3], 5.0)
        ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, bands, gains, seams)
    else:
        sharp = float(np.float32(1.0) / np.float32(bw))
        ref = oracle_loop_with_blender(rig, imgs, gains, seams, orc.SimpleBlender(1 if tag == "feather" else 0, sharp))
    assert tuple(out["dst_roi"]) == tuple(ref["dst_roi"])
    assert np.array_equal(out["mask"], ref["mask"]) and np.array_equal(out["result16"], ref["result16"])
    assert np.array_equal(out["result8"], np.clip(ref["result16"], 0, 255).astype(np.uint8))
    # a tiny blend_strength makes blend_width < 1: every type falls back to Blender::NO (image_stitching.cpp:1178-1179)
    c0 = isb.Composer(rig.warp, rig.scale, 3, blend_type=tag, blend_strength=1e-4)
    c0.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
    out0 = c0.run(imgs, gains, seams, want16=True)
    ref0 = oracle_loop_with_blender(rig, imgs, gains, seams, orc.SimpleBlender(0, 0.02))
    assert np.array_equal(out0["mask"], ref0["mask"]) and np.array_equal(out0["result16"], ref0["result16"])


def test_seam_aware_culling_follows_the_masks_of_each_run():
    """The per-run need map must track the seam masks actually passed: same cached plan, different masks (incl. all-zero for
    one image, none at all, and masks that keep only a corner), every run bit-exact against the oracle."""
    rig, imgs, gains, nb = make_case("cfg2", 8, 5)
    base = seam_masks_oracle(rig)
    c = isb.Composer(rig.warp, rig.scale, nb, cache_plan=True)
    c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
    variants = []
    variants.append(base)
    v = [m.copy() for m in base]          # image 0 contributes nothing, image 1 keeps only its top-left corner,
    v[0][:] = 0                           # image 2 takes everything it can see
    v[1][:] = 0
    v[1][: v[1].shape[0] // 3, : v[1].shape[1] // 3] = 255
    v[2][:] = 255
    variants.append(v)
    variants.append(None)                 # no seam masks at all after runs with masks
    variants.append([np.roll(m, m.shape[1] // 3, axis=1) for m in base])  # shifted support
    variants.append(base)
    for k, seams in enumerate(variants):
        out = c.run(imgs, gains, seams, want16=True)
        ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
        assert np.array_equal(out["mask"], ref["mask"]), k
        assert np.array_equal(out["result16"], ref["result16"]), k
        assert np.array_equal(out["result8"], ref["result8"]), k


@pytest.mark.parametrize("ttype", [0, 1])
def test_timelapser_bit_exact(ttype):
    """cv::detail::Timelapser / TimelapserCrop through the C ABI against the (cv2-pinned) numpy restatement."""
    rng = np.random.default_rng(12)
    for corners, sizes in [([(0, 0), (150, -30), (-77, 41)], [(300, 200), (257, 213), (190, 260)]),
                           ([(0, 0), (20, 10), (-15, 25)], [(300, 200), (257, 213), (290, 160)])]:
        a, b = isb.Timelapser_createDefault(ttype), orc.Timelapser(ttype)
        a.initialize(corners, sizes)
        b.initialize(corners, sizes)
        assert tuple(a.dst_roi) == tuple(b.roi)
        for (c, (sw, sh)) in zip(corners, sizes):
            img = rng.integers(-300, 600, (sh, sw, 3)).astype(np.int16)
            a.process(img, None, c)
            b.process(img, None, c)
            assert np.array_equal(a.getDst(), b.getDst()), (ttype, c)
    with pytest.raises(isb.IsbError):
        isb.Timelapser(0).process(np.zeros((4, 4, 3), np.int16), None, (0, 0))  # initialize() first

"""CPU: the C oracle (oracle/isb_oracle.c) against the cv2-generated golden vectors in tests/golden/."""
import os

import numpy as np
import pytest
from conftest import make_case, seam_masks_oracle

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {
    "cfg2_d16_nb3": ("cfg2", 16, 3, None, "texture"),
    "cfg2_d16_nb5_checker": ("cfg2", 16, 5, None, "checker"),
    "cfg4_d8_nb5": ("cfg4", 8, 5, None, "texture"),
    "cfg3_d32_nb4": ("cfg3", 32, 4, None, "texture"),
}


@pytest.mark.parametrize("name", list(CASES))
def test_compose_matches_golden(name):
    rigname, div, nb, mx, kind = CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    rig, imgs, gains, nb = make_case(rigname, div, nb, mx, kind)
    seams = seam_masks_oracle(rig)
    # the oracle's own nearest warp reproduces the cv2 seam masks
    assert [s.shape for s in seams] == [tuple(v) for v in g["seam_sizes"]]
    assert [int(s.astype(np.int64).sum()) for s in seams] == [int(v) for v in g["seam_sums"]]
    assert np.array_equal(seams[0], g["seam0"])
    out = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    assert out["corners"] == [tuple(v) for v in g["corners"]]
    assert out["sizes"] == [tuple(v) for v in g["sizes"]]
    assert out["dst_roi"] == tuple(g["dst_roi"])
    assert np.array_equal(out["mask"], g["mask"])
    assert np.array_equal(out["result16"], g["result16"])  # bit-exact, int16


def test_stage_vectors():
    g = np.load(os.path.join(GOLD, "cfg2_d16_nb3.npz"))
    rig, imgs, gains, nb = make_case("cfg2", 16, 3)
    _, iw = orc.warp(rig.warp, rig.scale, imgs[0], rig.Ks[0], rig.Rs[0], orc.LINEAR, 1)
    _, mw = orc.warp(rig.warp, rig.scale, np.full(imgs[0].shape[:2], 255, np.uint8), rig.Ks[0], rig.Rs[0], orc.NEAREST, 0)
    assert np.array_equal(mw, g["valid0"])
    assert np.array_equal(orc.gain_apply(iw, gains[0]), g["warped0"])
    up = orc.resize_linear_exact(orc.dilate3x3(g["seam0"]), mw.shape[1], mw.shape[0])
    assert np.array_equal(up & mw, g["mask0"])


def test_primitives():
    p = np.load(os.path.join(GOLD, "primitives.npz"))
    for i in range(3):
        assert np.array_equal(orc.pyrdown_16s(p[f"p16_{i}"]), p[f"down16_{i}"])
        assert np.array_equal(orc.pyrup_16s(p[f"p16_{i}"]), p[f"up16_{i}"])
        assert np.array_equal(orc.pyrdown_32f(p[f"w_{i}"]), p[f"downw_{i}"])  # float op order, bit-exact
    assert np.array_equal(orc.dilate3x3(p["mask"]), p["mask_dil"])
    assert np.array_equal(orc.resize_linear_exact(p["mask_dil"], 453, 260), p["mask_up"])
    gu = orc.resize_linear_f32(p["gain"], 451, 353)
    # the float gain-map upsample is the one step that is only pinned to <= 1 ulp
    assert np.max(np.abs(gu.view(np.int32).astype(np.int64) - p["gain_up"].view(np.int32))) <= 1
    assert np.mean(gu != p["gain_up"]) < 0.02


def simple_blend_inputs():
    """The seeded inputs tests/golden/make_golden.py::simple_blend_inputs used (kept in step with it)."""
    rng = np.random.default_rng(77)
    corners = [(0, 0), (150, -30), (-77, 41)]
    sizes = [(300, 200), (257, 213), (190, 260)]
    imgs, masks = [], []
    for (sw, sh) in sizes:
        imgs.append(rng.integers(0, 256, (sh, sw, 3)).astype(np.int16))
        m = np.zeros((sh, sw), np.uint8)
        m[10:-10, 10:-10] = 255
        m[20:40, 20:60] = rng.integers(0, 256, (20, 40))
        m[50:60, 50:90] = 0
        masks.append(m)
    return corners, sizes, imgs, masks


def test_simple_blenders_match_golden():
    """Blender::NO / FeatherBlender / createWeightMap restatements against cv2-generated vectors."""
    g = np.load(os.path.join(GOLD, "simple_blend.npz"))
    corners, sizes, imgs, masks = simple_blend_inputs()
    assert orc.result_roi(corners, sizes) == tuple(g["roi"])
    assert np.array_equal(orc.create_weight_map(masks[0], 0.02), g["wm0"])
    assert np.array_equal(orc.create_weight_map(np.full((40, 50), 255, np.uint8), 0.02), g["wm_full"])
    for tag, btype, sharp in [("no", 0, 0.02), ("feather", 1, 0.02), ("feather_sharp", 1, 1 / 37.3)]:
        b = orc.SimpleBlender(btype, sharp)
        b.prepare(tuple(g["roi"]))
        for img, m, c in zip(imgs, masks, corners):
            b.feed(img, m, c)
        r, rm = b.blend()
        assert np.array_equal(r, g[tag + "_result16"]) and np.array_equal(rm, g[tag + "_mask"]), tag


def oracle_loop_with_blender(rig, imgs, gains, seams, blender):
    """The compositing loop composed from the oracle's own functions for a prepare / feed / blend blender."""
    corners, sizes = [], []
    for K, R in zip(rig.Ks, rig.Rs):
        x, y, w, h = orc.warp_roi(rig.warp, rig.scale, rig.W, rig.H, K, R)
        corners.append((x, y))
        sizes.append((w, h))
    roi = orc.result_roi(corners, sizes)
    blender.prepare(roi)
    for i, (im, K, R) in enumerate(zip(imgs, rig.Ks, rig.Rs)):
        _, iw = orc.warp(rig.warp, rig.scale, im, K, R, orc.LINEAR, 1)
        _, mw = orc.warp(rig.warp, rig.scale, np.full(im.shape[:2], 255, np.uint8), K, R, orc.NEAREST, 0)
        iw = orc.gain_apply(iw, gains[i])
        up = orc.resize_linear_exact(orc.dilate3x3(seams[i]), mw.shape[1], mw.shape[0])
        blender.feed(iw.astype(np.int16), up & mw, corners[i])
    r, m = blender.blend()
    return dict(dst_roi=roi, result16=r, mask=m, result8=np.clip(r, 0, 255).astype(np.uint8))


@pytest.mark.parametrize("tag", ["feather", "no"])
def test_loop_with_simple_blenders_matches_golden(tag):
    """image_stitching.cpp:1086-1229 with blend_type feather / no, against the cv2-generated loop vectors.  The float
    gain-map upsample is pinned to <= 1 ulp only, hence the <= 1 grey-level bar instead of equality."""
    g = np.load(os.path.join(GOLD, "simple_blend.npz"))
    rig, imgs, gains, nb = make_case("cfg2", 16, 3)
    seams = seam_masks_oracle(rig)
    b = orc.SimpleBlender(1 if tag == "feather" else 0, float(g["loop_sharpness"]))
    out = oracle_loop_with_blender(rig, imgs, gains, seams, b)
    assert out["dst_roi"] == tuple(g["loop_dst_roi"])
    assert np.array_equal(out["mask"], g["loop_" + tag + "_mask"])
    d = np.abs(out["result16"].astype(int) - g["loop_" + tag + "_result16"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3

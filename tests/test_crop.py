"""crop() of the reference (cropper.cpp:116-209).

Checkers, strongest first:
 * oracle/_ref: the reference's OWN cropper.cpp, compiled unmodified; its two OpenCV contour calls are answered by the real
   OpenCV (cv2) through callbacks (oracle/cvshim/opencv2/imgproc.hpp), so crop() itself is the reference's code;
 * oracle/crop_oracle.py: numpy restatement in the parallel formulation the CUDA code uses (pinned against the former here).
CPU tests pin the restatement (and the shim's cvtColor) against cv2 / the compiled reference; GPU tests compare
isb_crop_rect / isb_crop_rect_image with both, bit for bit, on random masks, masks with holes, panorama masks of the rigs."""
import numpy as np
import pytest

from oracle import crop_oracle as co
from oracle import ref_helpers as ref

cv2 = pytest.importorskip("cv2")
have_ref = ref.available()
needs_ref = pytest.mark.skipif(not have_ref, reason="oracle/_ref/libisb_ref.so not built (no /root/reference here)")


def rand_mask(rng, h, w, kind):
    from scipy import ndimage as ndi
    if kind == 0:
        return (rng.random((h, w)) < 0.6).astype(np.uint8) * 255
    if kind == 1:
        m = np.zeros((h, w), np.uint8)
        for _ in range(6):
            cv2.circle(m, (int(rng.integers(0, w)), int(rng.integers(0, h))), int(rng.integers(2, 14)), 255, -1)
        for _ in range(4):
            cv2.circle(m, (int(rng.integers(0, w)), int(rng.integers(0, h))), int(rng.integers(1, 4)), 0, -1)
        return m
    if kind == 2:
        return (ndi.gaussian_filter(rng.random((h, w)), 2.5) > 0.48).astype(np.uint8) * 255
    m = np.zeros((h, w), np.uint8)  # a panorama-like blob with a few holes and a ragged border
    m[2:h - 3, 3:w - 2] = 255
    for _ in range(5):
        m[rng.integers(0, h), rng.integers(0, w)] = 0
    for _ in range(3):
        x = int(rng.integers(0, w))
        m[: int(rng.integers(1, max(2, h // 3))), x:x + int(rng.integers(1, 6))] = 0
    return m


def cases(seed, n, lo=6, hi=70):
    rng = np.random.default_rng(seed)
    for it in range(n):
        h, w = int(rng.integers(lo, hi)), int(rng.integers(lo, hi))
        m = rand_mask(rng, h, w, it % 4)
        if m.any():
            yield m


def degenerate(mask, rect):
    # rectangles that make the reference read outside its buffer (undefined there): a view of zero width in column 0 row 0
    x, y, w, h = rect
    return (w == 0 and x == 0 and y == 0) or (h == 0 and y == 0)


@needs_ref
def test_restatement_matches_reference_crop():
    ref.install_cv2_contours()
    n = 0
    for m in cases(3, 400):
        r_ref = ref.crop_rect(np.repeat(m[:, :, None], 3, 2))
        if degenerate(m, r_ref):
            continue
        assert co.crop_rect_from_mask(m) == r_ref
        n += 1
    assert n > 300


@needs_ref
def test_reference_crop_from_image_and_16s():
    """crop() takes the IMAGE (gray > 0): coloured 8UC3 and 16SC3 sources; the shim's cvtColor is OpenCV's."""
    ref.install_cv2_contours()
    rng = np.random.default_rng(4)
    for m in cases(5, 40, 12, 50):
        img = (rng.integers(0, 3, m.shape + (3,)) * (m[:, :, None] > 0)).astype(np.uint8)  # dark pixels: gray may round to 0
        g = co.gray_positive(img)
        assert np.array_equal(g, (cv2.cvtColor(img, cv2.COLOR_RGB2GRAY) > 0).astype(np.uint8) * 255)
        if not g.any():
            continue
        r = ref.crop_rect(img)
        if not degenerate(g, r):
            assert co.crop_rect_from_mask(g) == r
        img16 = img.astype(np.int16) * 3 - 2  # negative and > 0 values: saturate_cast to 8U first
        g16 = co.gray_positive(img16)
        if g16.any():
            r16 = ref.crop_rect(img16)
            if not degenerate(g16, r16):
                assert co.crop_rect_from_mask(g16) == r16


def test_contour_rules_match_cv2():
    """The three rules the parallel formulation rests on, against cv2.findContours / drawContours directly."""
    from scipy import ndimage as ndi
    for m in cases(6, 150):
        fg = m > 0
        cs, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        lab, l, v = co.choose_component(fg)
        # (1) per-pixel passage counts of every external contour
        got = np.zeros(m.shape, np.int64)
        for c in cs:
            for x, y in c.reshape(-1, 2):
                got[y, x] += 1
        assert np.array_equal(got, v)
        # (2) the contour crop() keeps: first maximum in cv's order
        k = int(np.argmax([len(c) for c in cs]))
        assert lab[cs[k][0, 0, 1], cs[k][0, 0, 0]] == l
        # (3) the filled contour
        cm = np.zeros_like(m)
        cv2.drawContours(cm, cs, k, 255, -1, 8)
        assert np.array_equal(cm > 0, ~co.outside_region(lab == l))


# ---- GPU --------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_crop_random_masks():
    import image_stitching_b200 as isb
    if have_ref:
        ref.install_cv2_contours()
    n = 0
    for m in cases(7, 300, 6, 90):
        want = co.crop_rect_from_mask(m)
        if have_ref:
            r_ref = ref.crop_rect(np.repeat(m[:, :, None], 3, 2))
            if degenerate(m, r_ref):
                continue
            assert want == r_ref
        rect, npts = isb.crop_rect(m, with_points=True)
        assert rect == want, (m.shape, rect, want)
        cs, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        assert npts == max(len(c) for c in cs)
        n += 1
    assert n > 200
    with pytest.raises(isb.IsbError):
        isb.crop_rect(np.zeros((20, 30), np.uint8))


@pytest.mark.gpu
def test_gpu_crop_panorama_masks_and_image_form():
    """Panorama masks of the rigs (wrap-around gaps, ragged top / bottom), a mask with holes, device pointers, and the
    literal crop(source) on the 16SC3 panorama."""
    import torch

    import image_stitching_b200 as isb
    from conftest import make_case, seam_masks_oracle
    if have_ref:
        ref.install_cv2_contours()
    for name, div, nb in (("cfg2", 8, 4), ("cfg4", 4, 5), ("cfg3", 16, 3)):
        rig, imgs, gains, nb = make_case(name, div, nb)
        seams = seam_masks_oracle(rig)
        if name == "cfg4":
            seams[1][10:40, 20:60] = 0  # punch a hole into the blend weights -> a hole in the result mask
        out = isb.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
        m = out["mask"]
        want = co.crop_rect_from_mask(m)
        if have_ref:
            assert want == ref.crop_rect(np.repeat(m[:, :, None], 3, 2))
        assert isb.crop_rect(m) == want
        assert isb.crop_rect(torch.from_numpy(m).cuda()) == want
        # crop(source) on what blend() returns (16SC3): its own mask is gray > 0 of the saturated image
        g = co.gray_positive(out["result16"])
        want_img = co.crop_rect_from_mask(g)
        if have_ref:
            assert want_img == ref.crop_rect(out["result16"])
        view, rect = isb.crop(out["result16"])
        assert rect == want_img and view.shape[:2] == (want_img[3], want_img[2])
        x, y, w, h = rect
        assert np.array_equal(view, np.clip(out["result16"], 0, 255).astype(np.uint8)[y:y + h, x:x + w])

"""TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's crop() (image_stitching/cropper.cpp:116-209, helper
checkInteriorExterior :6-104) in the PARALLEL formulation the CUDA implementation uses, so that formulation can be checked
against the reference's own code (oracle/_ref: cropper.cpp compiled unmodified, with cv2 answering findContours/drawContours)
on a machine without a GPU.

crop() only uses (a) which external contour has the most points, (b) the sorted x and the sorted y VALUES of that contour's
points (with multiplicity) and (c) the filled contour.  None of these needs the border-following trace itself:
 * an external contour exists per 8-connected component that touches the OUTSIDE background (the 4-connected background region
   connected to the image frame);
 * cv::findContours(CHAIN_APPROX_NONE) emits pixel p once per passage of the border walk, and p is passed once per maximal
   circular run of background pixels in its 8-ring that contains a 4-neighbour and belongs to the outside region (1 for an
   isolated pixel) - cropper.cpp:141-148 compares the sums of these multiplicities;
 * ties: findContours lists external contours in reverse raster order of their first pixel and crop() keeps the first maximum;
 * drawContours(filled) of the chosen contour = everything that a 4-connected flood from the frame through pixels NOT of that
   component cannot reach (cropper.cpp:153).
Each rule is verified against cv2 / the compiled reference in tests/test_crop.py."""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

RING = [(-1, -1), (-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1)]  # circular; odd positions are 4-neighbours


def gray_positive(img):
    """mask = cvtColor(convertTo(img, CV_8U), RGB2GRAY) > 0  (cropper.cpp:118-124); img 8UC3 or 16SC3."""
    a = np.clip(img, 0, 255).astype(np.int64)
    g = (a[..., 0] * 4899 + a[..., 1] * 9617 + a[..., 2] * 1868 + (1 << 13)) >> 14
    return (g > 0).astype(np.uint8) * 255


def outside_region(blocked):
    """4-connected flood from the (virtual, one-pixel) frame through pixels where `blocked` is False."""
    pad = np.pad(~blocked, 1, constant_values=True)
    lab, _ = ndi.label(pad)
    return (lab == lab[0, 0])[1:-1, 1:-1]


def visit_counts(fg, outside):
    """How often the external border walk passes every foreground pixel."""
    H, W = fg.shape
    fgp = np.pad(fg, 1, constant_values=False)
    outp = np.pad(outside, 1, constant_values=True)
    b = [~fgp[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] for dy, dx in RING]
    o = [outp[1 + dy:1 + dy + H, 1 + dx:1 + dx + W] for dy, dx in RING]
    v = np.zeros((H, W), np.int64)
    for j in (1, 3, 5, 7):
        v += b[j] & o[j] & ~(b[(j - 1) % 8] & b[(j - 2) % 8])
    allbg = np.logical_and.reduce(b)
    v = np.where(allbg, o[1].astype(np.int64), v)
    return np.where(fg, v, 0)


def choose_component(fg):
    """(label image, chosen label, visit counts) - the contour crop() keeps (cropper.cpp:141-148), or label 0 if none."""
    lab, n = ndi.label(fg, structure=np.ones((3, 3), int))
    out = outside_region(fg)
    v = visit_counts(fg, out)
    if n == 0:
        return lab, 0, v
    tot = ndi.sum(v, lab, index=np.arange(1, n + 1)).astype(np.int64)
    first = ndi.minimum(np.arange(fg.size).reshape(fg.shape), lab, index=np.arange(1, n + 1))  # raster-first pixel
    best = max(range(n), key=lambda k: (tot[k], first[k]))  # most points; ties: last in raster order (listed first by cv)
    if tot[best] == 0:
        return lab, 0, v
    return lab, best + 1, v


def crop_rect_from_mask(mask):
    """Rectangle (x, y, w, h) crop() narrows the image to, for the binary mask `mask != 0`."""
    fg = np.asarray(mask) != 0
    H, W = fg.shape
    lab, l, v = choose_component(fg)
    if l == 0:
        raise ValueError("no contour (cropper.cpp:151 would throw on contours.at(0))")
    C = lab == l
    filled = ~outside_region(C)
    vc = np.where(C, v, 0)
    sx = np.repeat(np.arange(W), vc.sum(axis=0))  # sorted x values of the contour points
    sy = np.repeat(np.arange(H), vc.sum(axis=1))
    flat = filled.reshape(-1)

    def at(y, x):  # Mat::at on the ROI view: plain pointer arithmetic on the continuous parent buffer
        i = y * W + x
        return bool(flat[i]) if 0 <= i < flat.size else False

    a, b, c, d = 0, len(sx) - 1, 0, len(sy) - 1
    rect = (0, 0, 0, 0)
    while a < b and c < d:
        rx, ry, rw, rh = int(sx[a]), int(sy[c]), int(sx[b] - sx[a]), int(sy[d] - sy[c])
        rect = (rx, ry, rw, rh)
        top = sum(not at(ry, rx + x) for x in range(rw))
        bottom = sum(not at(ry + rh - 1, rx + x) for x in range(rw))
        left = sum(not at(ry + y, rx) for y in range(rh))
        right = sum(not at(ry + y, rx + rw - 1) for y in range(rh))
        if top == 0 and bottom == 0 and left == 0 and right == 0:
            break
        oc_t = oc_b = oc_l = oc_r = 0
        if top > bottom:
            if top > left and top > right:
                oc_t = 1
        elif bottom > left:
            if bottom > right:
                oc_b = 1
        if left >= right:
            if left >= bottom and left >= top:
                oc_l = 1
        elif right >= top:
            if right >= bottom:
                oc_r = 1
        a += oc_l
        b -= oc_r
        c += oc_t
        d -= oc_b
    return rect

"""ctypes binding of oracle/libisb_oracle.so (TEST INFRASTRUCTURE ONLY - see isb_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
KIND = {"spherical": 0, "cylindrical": 1}
NEAREST, LINEAR = 0, 1


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libisb_oracle.so")
    src = os.path.join(_HERE, "isb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libisb_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_blender_create.restype = C.c_void_p
        _LIB.orc_simple_blender_create.restype = C.c_void_p
        _LIB.orc_simple_blender_create.argtypes = [C.c_int, C.c_float]
        _LIB.orc_blender_num_bands.restype = C.c_int
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def warp_roi(kind, scale, w, h, K, R):
    r = np.zeros(4, np.int32)
    K, R = _f32(K), _f32(R)
    lib().orc_warp_roi(KIND[kind], C.c_float(scale), int(w), int(h), _p(K), _p(R), _p(r))
    return tuple(int(v) for v in r)


def build_maps(kind, scale, w, h, K, R):
    x, y, rw, rh = warp_roi(kind, scale, w, h, K, R)
    xm = np.empty((rh, rw), np.float32)
    ym = np.empty((rh, rw), np.float32)
    K, R = _f32(K), _f32(R)
    lib().orc_build_maps(KIND[kind], C.c_float(scale), int(w), int(h), _p(K), _p(R), _p(xm), _p(ym))
    return (x, y, rw, rh), xm, ym


def map_forward(kind, scale, K, R, x, y):
    o = np.zeros(2, np.float32)
    K, R = _f32(K), _f32(R)
    lib().orc_map_forward(KIND[kind], C.c_float(scale), _p(K), _p(R), C.c_float(x), C.c_float(y), _p(o))
    return float(o[0]), float(o[1])


def map_backward(kind, scale, K, R, u, v):
    o = np.zeros(2, np.float32)
    K, R = _f32(K), _f32(R)
    lib().orc_map_backward(KIND[kind], C.c_float(scale), _p(K), _p(R), C.c_float(u), C.c_float(v), _p(o))
    return float(o[0]), float(o[1])


def remap(src, xmap, ymap, interp, border):
    src = np.ascontiguousarray(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    dh, dw = xmap.shape
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, ch), np.uint8)
    xm, ym = _f32(xmap), _f32(ymap)
    fn = lib().orc_remap_linear_8u if interp == LINEAR else lib().orc_remap_nearest_8u
    fn(_p(src), w, h, ch, C.c_size_t(w * ch), _p(xm), _p(ym), dw, dh, _p(dst), C.c_size_t(dw * ch), int(border))
    return dst


def warp(kind, scale, src, K, R, interp, border):
    src = np.ascontiguousarray(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    x, y, rw, rh = warp_roi(kind, scale, w, h, K, R)
    dst = np.empty((rh, rw) if src.ndim == 2 else (rh, rw, ch), np.uint8)
    c = np.zeros(2, np.int32)
    K, R = _f32(K), _f32(R)
    lib().orc_warp(KIND[kind], C.c_float(scale), _p(src), w, h, ch, C.c_size_t(w * ch), _p(K), _p(R), int(interp),
                   int(border), _p(dst), C.c_size_t(rw * ch), _p(c))
    return (int(c[0]), int(c[1])), dst


def warp_backward(kind, scale, src, K, R, interp, border, dst_size):
    """RotationWarper::warpBackward: src = warped image (warpRoi(dst_size) large), returns the dst_size = (w, h) frame."""
    src = np.ascontiguousarray(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    dw, dh = dst_size
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, ch), np.uint8)
    K, R = _f32(K), _f32(R)
    rc = lib().orc_warp_backward(KIND[kind], C.c_float(scale), _p(src), w, h, ch, C.c_size_t(w * ch), _p(K), _p(R), int(interp),
                                 int(border), int(dw), int(dh), _p(dst), C.c_size_t(dw * ch))
    if rc != 0:
        raise ValueError("src is not warpRoi(dst_size) large")
    return dst


def dilate3x3(m):
    m = np.ascontiguousarray(m, np.uint8)
    o = np.empty_like(m)
    lib().orc_dilate3x3_8u(_p(m), m.shape[1], m.shape[0], _p(o))
    return o


def resize_linear_exact(m, dw, dh):
    m = np.ascontiguousarray(m, np.uint8)
    o = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_exact_8u(_p(m), m.shape[1], m.shape[0], _p(o), int(dw), int(dh))
    return o


def resize_linear_exact_ex(img, dw, dh, fx=0.0, fy=0.0):
    img = np.ascontiguousarray(img, np.uint8)
    ch = 1 if img.ndim == 2 else img.shape[2]
    o = np.empty((dh, dw) if img.ndim == 2 else (dh, dw, ch), np.uint8)
    lib().orc_resize_linear_exact_8u_ex(_p(img), img.shape[1], img.shape[0], ch, _p(o), int(dw), int(dh), C.c_double(fx),
                                        C.c_double(fy))
    return o


def rotate(img, code):
    """code 0 = ROTATE_90_CLOCKWISE, 1 = ROTATE_180"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    shape = (w, h) if code == 0 else (h, w)
    o = np.empty(shape if img.ndim == 2 else shape + (ch,), np.uint8)
    lib().orc_rotate_8u(_p(img), w, h, ch, int(code), _p(o))
    return o


def resize_linear_f32(g, dw, dh):
    g = _f32(g)
    o = np.empty((dh, dw), np.float32)
    lib().orc_resize_linear_f32(_p(g), g.shape[1], g.shape[0], _p(o), int(dw), int(dh))
    return o


def gain_apply(img, gain):
    img = np.ascontiguousarray(img, np.uint8).copy()
    g = _f32(gain)
    h, w = img.shape[:2]
    lib().orc_gain_apply_8uc3(_p(img), w, h, C.c_size_t(w * 3), _p(g), g.shape[1], g.shape[0])
    return img


def pyrdown_16s(a):
    a = np.ascontiguousarray(a, np.int16)
    ch = 1 if a.ndim == 2 else a.shape[2]
    h, w = a.shape[:2]
    o = np.empty(((h + 1) // 2, (w + 1) // 2) + (() if a.ndim == 2 else (ch,)), np.int16)
    lib().orc_pyrdown_16s(_p(a), w, h, ch, _p(o))
    return o


def pyrup_16s(a):
    a = np.ascontiguousarray(a, np.int16)
    ch = 1 if a.ndim == 2 else a.shape[2]
    h, w = a.shape[:2]
    o = np.empty((2 * h, 2 * w) + (() if a.ndim == 2 else (ch,)), np.int16)
    lib().orc_pyrup_16s(_p(a), w, h, ch, _p(o))
    return o


def pyrdown_32f(a):
    a = _f32(a)
    h, w = a.shape
    o = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    lib().orc_pyrdown_32f(_p(a), w, h, _p(o))
    return o


def result_roi(corners, sizes):
    c = np.ascontiguousarray(corners, np.int32).reshape(-1, 2)
    s = np.ascontiguousarray(sizes, np.int32).reshape(-1, 2)
    r = np.zeros(4, np.int32)
    lib().orc_result_roi(_p(c), _p(s), len(c), _p(r))
    return tuple(int(v) for v in r)


class Blender:
    """MultiBandBlender restatement with the cv2 call surface (prepare / feed / blend)."""

    def __init__(self, num_bands=5):
        self._h = C.c_void_p(lib().orc_blender_create(int(num_bands)))

    def __del__(self):
        try:
            lib().orc_blender_destroy(self._h)
        except Exception:
            pass

    def prepare(self, roi):
        r = np.ascontiguousarray(roi, np.int32)
        lib().orc_blender_prepare(self._h, _p(r))

    def numBands(self):
        return lib().orc_blender_num_bands(self._h)

    def rois(self):
        a, b = np.zeros(4, np.int32), np.zeros(4, np.int32)
        lib().orc_blender_get_rois(self._h, _p(a), _p(b))
        return tuple(int(v) for v in a), tuple(int(v) for v in b)

    def tile_rect(self, w, h, tl):
        a, b = np.zeros(2, np.int32), np.zeros(2, np.int32)
        lib().orc_blender_tile_rect(self._h, int(w), int(h), int(tl[0]), int(tl[1]), _p(a), _p(b))
        return (int(a[0]), int(a[1]), int(b[0]), int(b[1]))

    def feed(self, img16, mask, tl):
        img16 = np.ascontiguousarray(img16, np.int16)
        mask = np.ascontiguousarray(mask, np.uint8)
        h, w = mask.shape
        lib().orc_blender_feed(self._h, _p(img16), _p(mask), w, h, int(tl[0]), int(tl[1]))

    def blend(self):
        _, rf = self.rois()
        dst = np.empty((rf[3], rf[2], 3), np.int16)
        m = np.empty((rf[3], rf[2]), np.uint8)
        lib().orc_blender_blend(self._h, _p(dst), _p(m))
        return dst, m


def create_weight_map(mask, sharpness):
    mask = np.ascontiguousarray(mask, np.uint8)
    o = np.empty(mask.shape, np.float32)
    lib().orc_create_weight_map(_p(mask), mask.shape[1], mask.shape[0], C.c_float(sharpness), _p(o))
    return o


class SimpleBlender:
    """Blender::NO (type 0) and FeatherBlender (type 1) restatements with the cv2 call surface."""

    def __init__(self, btype, sharpness=0.02):
        self._h = C.c_void_p(lib().orc_simple_blender_create(int(btype), float(sharpness)))

    def __del__(self):
        try:
            lib().orc_simple_blender_destroy(self._h)
        except Exception:
            pass

    def prepare(self, roi):
        self.roi = tuple(int(v) for v in roi)
        r = np.ascontiguousarray(roi, np.int32)
        lib().orc_simple_blender_prepare(self._h, _p(r))

    def feed(self, img16, mask, tl):
        img16 = np.ascontiguousarray(img16, np.int16)
        mask = np.ascontiguousarray(mask, np.uint8)
        h, w = mask.shape
        lib().orc_simple_blender_feed(self._h, _p(img16), _p(mask), w, h, int(tl[0]), int(tl[1]))

    def blend(self):
        dst = np.empty((self.roi[3], self.roi[2], 3), np.int16)
        m = np.empty((self.roi[3], self.roi[2]), np.uint8)
        lib().orc_simple_blender_blend(self._h, _p(dst), _p(m))
        return dst, m


def compose(images, Ks, Rs, scale, kind, nb, gains=None, seam_masks=None):
    """Whole loop in C (image_stitching.cpp:1086-1229).  Returns dict like cv_reference.compose_cv."""
    n = len(images)
    imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
    wh = np.array([[im.shape[1], im.shape[0]] for im in imgs], np.int32)
    Kf = _f32(np.stack(Ks))
    Rf = _f32(np.stack(Rs))
    corners = np.zeros((n, 2), np.int32)
    sizes = np.zeros((n, 2), np.int32)
    roi = np.zeros(4, np.int32)
    lib().orc_compose_roi(KIND[kind], C.c_float(scale), n, _p(wh), _p(Kf), _p(Rf), _p(corners), _p(sizes), _p(roi))
    PA = C.c_void_p * n
    ip = PA(*[im.ctypes.data for im in imgs])
    gp = gwh = sp = swh = None
    if gains is not None:
        gs = [_f32(g) for g in gains]
        gp = PA(*[g.ctypes.data for g in gs])
        gwh = np.array([[g.shape[1], g.shape[0]] for g in gs], np.int32)
    if seam_masks is not None:
        ss = [np.ascontiguousarray(s, np.uint8) for s in seam_masks]
        sp = PA(*[s.ctypes.data for s in ss])
        swh = np.array([[s.shape[1], s.shape[0]] for s in ss], np.int32)
    out16 = np.empty((roi[3], roi[2], 3), np.int16)
    out8 = np.empty((roi[3], roi[2], 3), np.uint8)
    om = np.empty((roi[3], roi[2]), np.uint8)
    lib().orc_compose(KIND[kind], C.c_float(scale), n, ip, _p(wh), _p(Kf), _p(Rf), gp,
                      _p(gwh) if gwh is not None else None, sp, _p(swh) if swh is not None else None, int(nb),
                      _p(out16), _p(out8), _p(om))
    return dict(corners=[tuple(int(v) for v in c) for c in corners], sizes=[tuple(int(v) for v in s) for s in sizes],
                dst_roi=tuple(int(v) for v in roi), result16=out16, result8=out8, mask=om)


class Timelapser:
    """cv::detail::Timelapser (type 0, AS_IS) / TimelapserCrop (type 1) restated in numpy (image_stitching.cpp:1194-1215):
    process() clears the canvas and copies the pixels of one 16SC3 image that fall inside dst_roi."""

    def __init__(self, ttype):
        self.ttype = int(ttype)

    def initialize(self, corners, sizes):
        if self.ttype == 0:
            self.roi = result_roi(corners, sizes)
        else:  # Rect(Point(max tl), Point(min br)): cv::Rect_(pt1, pt2) orders the two corners itself
            tlx = max(c[0] for c in corners); tly = max(c[1] for c in corners)
            brx = min(c[0] + s[0] for c, s in zip(corners, sizes)); bry = min(c[1] + s[1] for c, s in zip(corners, sizes))
            self.roi = (min(tlx, brx), min(tly, bry), abs(brx - tlx), abs(bry - tly))
        self.dst = np.zeros((self.roi[3], self.roi[2], 3), np.int16)

    def process(self, img, mask, tl):
        self.dst[:] = 0
        h, w = img.shape[:2]
        x0, y0 = tl[0] - self.roi[0], tl[1] - self.roi[1]
        xa, ya = max(x0, 0), max(y0, 0)
        xb, yb = min(x0 + w, self.roi[2]), min(y0 + h, self.roi[3])
        if xb > xa and yb > ya:
            self.dst[ya:yb, xa:xb] = img[ya - y0:yb - y0, xa - x0:xb - x0]

    def getDst(self):
        return self.dst

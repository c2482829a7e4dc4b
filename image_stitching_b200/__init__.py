"""image_stitching_b200 - B200-native compositing path (rotation warp + multi-band blend).

This package is a thin ctypes mirror of the C ABI in include/image_stitching.h (libisb.so, built
in-tree by image_stitching_b200/build.py with nvcc for sm_100a).  The class and method names follow
the OpenCV objects the reference drives (image_stitching.cpp:1086-1229) so that code written against
`cv2.PyRotationWarper`, `cv2.detail_BlocksGainCompensator` and `cv2.detail_MultiBandBlender` reads the
same here.  There is NO CPU fallback: if libisb.so is missing or no CUDA device is present the calls
fail loudly.

Arrays may be numpy arrays (host) or torch CUDA tensors (device pointers are used in place).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ISB_LIBRARY: another build of the same library (A/B of compile-time tunables, tools/ab_env.py); never a fallback
_SO = os.environ.get("ISB_LIBRARY") or os.path.join(_HERE, "libisb.so")
_lib = None

SPHERICAL, CYLINDRICAL = 0, 1
INTER_NEAREST, INTER_LINEAR = 0, 1
BORDER_CONSTANT, BORDER_REFLECT = 0, 2
EULER = {"XYZ": 0, "YXZ": 1, "ZXY": 2, "ZYX": 3, "YZX": 4, "XZY": 5}
_KIND = {"spherical": SPHERICAL, "cylindrical": CYLINDRICAL, SPHERICAL: SPHERICAL, CYLINDRICAL: CYLINDRICAL}


class IsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class Camera(C.Structure):
    """cv::detail::CameraParams (isb_camera)."""
    _fields_ = [("focal", C.c_double), ("aspect", C.c_double), ("ppx", C.c_double), ("ppy", C.c_double),
                ("R", C.c_float * 9), ("t", C.c_float * 3)]

    @staticmethod
    def make(focal, ppx, ppy, R, aspect=1.0, t=(0, 0, 0)):
        c = Camera()
        c.focal, c.aspect, c.ppx, c.ppy = float(focal), float(aspect), float(ppx), float(ppy)
        c.R[:] = [float(v) for v in np.asarray(R, np.float32).reshape(9)]
        c.t[:] = [float(v) for v in t]
        return c

    def K(self):
        out = np.zeros(9, np.float32)
        lib().isb_camera_K(C.byref(self), out.ctypes.data_as(C.c_void_p))
        return out.reshape(3, 3)


class _Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("pitch", C.c_size_t)]


class _Gain(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int), ("height", C.c_int)]


class _Mask(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("pitch", C.c_size_t)]


class Config(C.Structure):
    _fields_ = [("warp_kind", C.c_int), ("warped_image_scale", C.c_float), ("num_bands", C.c_int),
                ("strip_index", C.c_int), ("strip_count", C.c_int), ("cache_plan", C.c_int), ("async_mode", C.c_int),
                ("gather_mode", C.c_int), ("pipeline_depth", C.c_int), ("use_blend_rule", C.c_int),
                ("blend_type", C.c_int), ("blend_strength", C.c_float), ("compose_scale", C.c_double), ("ingest_rotate", C.c_int),
                ("reserved", C.c_int * 3)]


class _Pano(C.Structure):
    _fields_ = [("data", C.c_void_p), ("pitch", C.c_size_t), ("mask", C.c_void_p), ("mask_pitch", C.c_size_t),
                ("data16", C.c_void_p), ("pitch16", C.c_size_t), ("roi_xywh", C.c_int * 4), ("strip_y0", C.c_int),
                ("strip_y1", C.c_int)]


def lib():
    """Loads libisb.so (building it when nvcc is available and it is missing).  Never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if "ISB_LIBRARY" not in os.environ:
        from . import build as _build
        if not _build.up_to_date():
            # sources newer than the binary (or no binary): rebuild when a compiler is here, never run a stale library silently
            try:
                _build.build()
            except RuntimeError:
                if not os.path.exists(_SO):
                    raise
                import warnings
                warnings.warn("libisb.so is older than csrc/ and nvcc is not available to rebuild it")
    L = C.CDLL(_SO)
    L.isb_last_error.restype = C.c_char_p
    L.isb_version.restype = C.c_char_p
    L.isb_launch_count.restype = C.c_longlong
    L.isb_composer_last_h2d_bytes.restype = C.c_longlong
    L.isb_composer_last_h2d_bytes.argtypes = [C.c_void_p]
    L.isb_composer_stage_name.restype = C.c_char_p
    L.isb_warper_get_scale.restype = C.c_float
    for f in ("isb_warper_create", "isb_compensator_create", "isb_blender_create", "isb_composer_create",
              "isb_simple_blender_create", "isb_timelapser_create"):
        getattr(L, f).restype = C.c_void_p
    L.isb_simple_blender_create.argtypes = [C.c_int, C.c_float]
    L.isb_simple_blender_set_sharpness.argtypes = [C.c_void_p, C.c_float]
    L.isb_simple_blender_sharpness.argtypes = [C.c_void_p]
    L.isb_simple_blender_sharpness.restype = C.c_float
    L.isb_create_weight_map.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_size_t]
    L.isb_warper_create.argtypes = [C.c_int, C.c_float]
    L.isb_warper_set_scale.argtypes = [C.c_void_p, C.c_float]
    L.isb_num_bands_for.argtypes = [C.c_int, C.c_int, C.c_float]
    L.isb_resize_linear_exact.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int,
                                          C.c_size_t, C.c_double, C.c_double]
    L.isb_quat_slerp.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    L.isb_quat_from_axis_angle.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    for f in ("isb_warper_destroy", "isb_compensator_destroy", "isb_blender_destroy", "isb_composer_destroy",
              "isb_simple_blender_destroy", "isb_timelapser_destroy"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = None
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise IsbError(rc, lib().isb_last_error().decode())


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class DevPtr:
    """A raw device pointer (own allocation or a peer buffer opened through CUDA IPC) with an array shape."""

    def __init__(self, ptr, shape, owner=False, ipc=False):
        self.ptr, self.shape, self._owner, self._ipc = int(ptr), tuple(shape), owner, ipc

    @staticmethod
    def alloc(shape, itemsize=1):
        n = int(np.prod(shape)) * itemsize
        p = C.c_void_p()
        _chk(lib().isb_device_malloc(C.c_size_t(n), C.byref(p)))
        return DevPtr(p.value, shape, owner=True)

    def nbytes(self, itemsize=1):
        return int(np.prod(self.shape)) * itemsize

    def ipc_handle(self):
        h = (C.c_ubyte * 64)()
        _chk(lib().isb_ipc_get_handle(C.c_void_p(self.ptr), h))
        return bytes(h)

    @staticmethod
    def open_ipc(handle, shape):
        h = (C.c_ubyte * 64)(*handle)
        p = C.c_void_p()
        _chk(lib().isb_ipc_open_handle(h, C.byref(p)))
        return DevPtr(p.value, shape, ipc=True)

    def to_numpy(self, dtype=np.uint8, out=None):
        out = np.empty(self.shape, dtype) if out is None else out
        _chk(lib().isb_memcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), C.c_size_t(out.nbytes), 1))
        return out

    def close(self):
        if self.ptr:
            if self._ipc:
                lib().isb_ipc_close_handle(C.c_void_p(self.ptr))
            elif self._owner:
                lib().isb_device_free(C.c_void_p(self.ptr))
            self.ptr = 0


def _ptr(a):
    """(address, keepalive) of a numpy array or torch tensor."""
    if a is None:
        return None, None
    if isinstance(a, DevPtr):
        return a.ptr, a
    if _is_torch(a):
        a = a if a.is_contiguous() else a.contiguous()
        return a.data_ptr(), a
    a = np.ascontiguousarray(a)
    return a.ctypes.data, a


def _f32p(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    return a.ctypes.data_as(C.c_void_p), a


def _shape(a):
    return tuple(int(v) for v in a.shape)


def device_count():
    return int(lib().isb_device_count())


def set_stream(cuda_stream_ptr):
    lib().isb_set_stream(C.c_void_p(int(cuda_stream_ptr) if cuda_stream_ptr else None))


def launch_count(reset=False):
    return int(lib().isb_launch_count(1 if reset else 0))


# ---------------------------------------------------------------------------------------------------
# pose helpers (quaternion.h / euler.h / serializer.cpp mirrors)
# ---------------------------------------------------------------------------------------------------
def _d(a, n):
    a = np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1))
    assert a.size == n
    return a


def quat_from_rotation_matrix(R):
    R, q = _d(R, 9), np.zeros(4)
    lib().isb_quat_from_rotation_matrix(R.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
    return q


def quat_to_rotation_matrix(q):
    q, R = _d(q, 4), np.zeros(9)
    lib().isb_quat_to_rotation_matrix(q.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
    return R.reshape(3, 3)


def quat_from_euler(e, order="XYZ"):
    e, q = _d(e, 3), np.zeros(4)
    lib().isb_quat_from_euler(e.ctypes.data_as(C.c_void_p), EULER[order], q.ctypes.data_as(C.c_void_p))
    return q


def quat_from_axis_angle(axis, angle):
    a, q = _d(axis, 3), np.zeros(4)
    lib().isb_quat_from_axis_angle(a.ctypes.data_as(C.c_void_p), float(angle), q.ctypes.data_as(C.c_void_p))
    return q


def quat_multiply(a, b):
    a, b, q = _d(a, 4), _d(b, 4), np.zeros(4)
    lib().isb_quat_multiply(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
    return q


def quat_slerp(a, b, t):
    a, b, q = _d(a, 4), _d(b, 4), np.zeros(4)
    lib().isb_quat_slerp(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), float(t),
                         q.ctypes.data_as(C.c_void_p))
    return q


def pose_from_cam_transform(R, is_portrait):
    R, o = _d(R, 9), np.zeros(9)
    lib().isb_pose_from_cam_transform(R.ctypes.data_as(C.c_void_p), int(bool(is_portrait)), o.ctypes.data_as(C.c_void_p))
    return o.reshape(3, 3)


def rotationMatrixToEulerAngles(R, order="XYZ"):
    R, e = _d(R, 9), np.zeros(3)
    _chk(lib().isb_rotation_matrix_to_euler(R.ctypes.data_as(C.c_void_p), EULER[order], e.ctypes.data_as(C.c_void_p)))
    return e


def eulerAnglesToRotationMatrix(e, order="XYZ"):
    e, R = _d(e, 3), np.zeros(9)
    _chk(lib().isb_euler_to_rotation_matrix(e.ctypes.data_as(C.c_void_p), EULER[order], R.ctypes.data_as(C.c_void_p)))
    return R.reshape(3, 3)


def parseMatrixStr(s):
    out = np.zeros(1024)
    side = C.c_int(0)
    _chk(lib().isb_parse_matrix_str(s.encode(), out.ctypes.data_as(C.c_void_p), out.size, C.byref(side)))
    return out[: side.value ** 2].reshape(side.value, side.value).copy()


def serializeMatrix(m):
    m = np.asarray(m)
    is32 = m.dtype == np.float32
    md = np.ascontiguousarray(m, np.float64)
    rows, cols = (md.shape + (1,))[:2] if md.ndim == 1 else md.shape
    buf = C.create_string_buffer(64 * md.size + 16)
    _chk(lib().isb_serialize_matrix(md.ctypes.data_as(C.c_void_p), int(rows), int(cols), int(is32), buf, len(buf)))
    return buf.value.decode()


def deserializeMatrix(s):
    out = np.zeros(4096, np.float32)
    r, c = C.c_int(0), C.c_int(0)
    _chk(lib().isb_deserialize_matrix(s.encode(), out.ctypes.data_as(C.c_void_p), out.size, C.byref(r), C.byref(c)))
    return out[: r.value * c.value].reshape(r.value, c.value).copy()


def serializeCameraParams(cams, path=None):
    arr = (Camera * len(cams))(*cams)
    _chk(lib().isb_save_cams(path.encode() if path else None, arr, len(cams)))


def deserializeCameraParams(path=None):
    n = C.c_int(0)
    _chk(lib().isb_load_cams(path.encode() if path else None, None, 0, C.byref(n)))
    arr = (Camera * max(n.value, 1))()
    _chk(lib().isb_load_cams(path.encode() if path else None, arr, n.value, C.byref(n)))
    return [arr[i] for i in range(n.value)]


def serializeIndices(idx, path=None):
    a = np.ascontiguousarray(idx, np.int32)
    _chk(lib().isb_save_indices(path.encode() if path else None, a.ctypes.data_as(C.c_void_p), a.size))


def deserializeIndices(path=None):
    n = C.c_int(0)
    _chk(lib().isb_load_indices(path.encode() if path else None, None, 0, C.byref(n)))
    a = np.zeros(max(n.value, 1), np.int32)
    _chk(lib().isb_load_indices(path.encode() if path else None, a.ctypes.data_as(C.c_void_p), a.size, C.byref(n)))
    return [int(v) for v in a[: n.value]]


# ---------------------------------------------------------------------------------------------------
# cv::detail::RotationWarper
# ---------------------------------------------------------------------------------------------------
class RotationWarper:
    """Mirror of cv2.PyRotationWarper(kind, scale) for 'spherical' and 'cylindrical'."""

    def __init__(self, kind, scale):
        if kind not in _KIND:
            raise IsbError(-5, f"unsupported warper type {kind!r}")
        self._h = C.c_void_p(lib().isb_warper_create(_KIND[kind], float(scale)))
        if not self._h:
            raise IsbError(-5, lib().isb_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().isb_warper_destroy(self._h)
            self._h = None

    def getScale(self):
        return float(lib().isb_warper_get_scale(self._h))

    def setScale(self, s):
        _chk(lib().isb_warper_set_scale(self._h, float(s)))

    def warpRoi(self, src_size, K, R):
        Kp, _k = _f32p(K)
        Rp, _r = _f32p(R)
        r = (C.c_int * 4)()
        _chk(lib().isb_warper_warp_roi(self._h, int(src_size[0]), int(src_size[1]), Kp, Rp, r))
        return tuple(r)

    def warpPoint(self, pt, K, R):
        Kp, _k = _f32p(K)
        Rp, _r = _f32p(R)
        p, _p = _f32p(pt)
        o = np.zeros(2, np.float32)
        _chk(lib().isb_warper_warp_point(self._h, p, Kp, Rp, o.ctypes.data_as(C.c_void_p)))
        return float(o[0]), float(o[1])

    def warpPointBackward(self, pt, K, R):
        Kp, _k = _f32p(K)
        Rp, _r = _f32p(R)
        p, _p = _f32p(pt)
        o = np.zeros(2, np.float32)
        _chk(lib().isb_warper_warp_point_backward(self._h, p, Kp, Rp, o.ctypes.data_as(C.c_void_p)))
        return float(o[0]), float(o[1])

    def buildMaps(self, src_size, K, R):
        x, y, w, h = self.warpRoi(src_size, K, R)
        Kp, _k = _f32p(K)
        Rp, _r = _f32p(R)
        xm, ym = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
        r = (C.c_int * 4)()
        _chk(lib().isb_warper_build_maps(self._h, int(src_size[0]), int(src_size[1]), Kp, Rp,
                                         xm.ctypes.data_as(C.c_void_p), ym.ctypes.data_as(C.c_void_p),
                                         C.c_size_t(w * 4), r))
        return tuple(r), xm, ym

    def warp(self, src, K, R, interp_mode, border_mode):
        src = np.ascontiguousarray(src, np.uint8)
        h, w = src.shape[:2]
        ch = 1 if src.ndim == 2 else src.shape[2]
        x, y, rw, rh = self.warpRoi((w, h), K, R)
        dst = np.empty((rh, rw) if src.ndim == 2 else (rh, rw, ch), np.uint8)
        Kp, _k = _f32p(K)
        Rp, _r = _f32p(R)
        c = (C.c_int * 2)()
        _chk(lib().isb_warper_warp(self._h, src.ctypes.data_as(C.c_void_p), w, h, ch, C.c_size_t(w * ch), Kp, Rp,
                                   int(interp_mode), int(border_mode), dst.ctypes.data_as(C.c_void_p),
                                   C.c_size_t(rw * ch), c))
        return (c[0], c[1]), dst


    def warpBackward(self, src, K, R, interp_mode, border_mode, dst_size):
        """cv2.PyRotationWarper.warpBackward: src = warped image of warpRoi(dst_size) pixels; returns the dst_size = (w, h) frame."""
        src = np.ascontiguousarray(src, np.uint8)
        h, w = src.shape[:2]
        ch = 1 if src.ndim == 2 else src.shape[2]
        dw, dh = int(dst_size[0]), int(dst_size[1])
        dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, ch), np.uint8)
        Kp, _k = _f32p(K)
        Rp, _r = _f32p(R)
        _chk(lib().isb_warper_warp_backward(self._h, src.ctypes.data_as(C.c_void_p), w, h, ch, C.c_size_t(w * ch), Kp, Rp,
                                            int(interp_mode), int(border_mode), dw, dh, dst.ctypes.data_as(C.c_void_p),
                                            C.c_size_t(dw * ch)))
        return dst


# ---------------------------------------------------------------------------------------------------
# cv::detail::BlocksGainCompensator (apply side)
# ---------------------------------------------------------------------------------------------------
class BlocksGainCompensator:
    def __init__(self, bl_width=32, bl_height=32, nr_feeds=1):
        self._h = C.c_void_p(lib().isb_compensator_create(int(bl_width), int(bl_height)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().isb_compensator_destroy(self._h)
            self._h = None

    def setMatGains(self, gains):
        gs = [np.ascontiguousarray(g, np.float32) for g in gains]
        n = len(gs)
        ptrs = (C.c_void_p * n)(*[g.ctypes.data for g in gs])
        gw = (C.c_int * n)(*[g.shape[1] for g in gs])
        gh = (C.c_int * n)(*[g.shape[0] for g in gs])
        _chk(lib().isb_compensator_set_mat_gains(self._h, n, ptrs, gw, gh))

    def getMatGain(self, index):
        gw, gh = C.c_int(0), C.c_int(0)
        _chk(lib().isb_compensator_get_mat_gain(self._h, int(index), None, 0, C.byref(gw), C.byref(gh)))
        out = np.zeros((gh.value, gw.value), np.float32)
        _chk(lib().isb_compensator_get_mat_gain(self._h, int(index), out.ctypes.data_as(C.c_void_p), out.size,
                                                C.byref(gw), C.byref(gh)))
        return out

    def apply(self, index, corner, image, mask=None):
        img = np.ascontiguousarray(image, np.uint8).copy()
        h, w = img.shape[:2]
        c = (C.c_int * 2)(int(corner[0]), int(corner[1]))
        _chk(lib().isb_compensator_apply(self._h, int(index), c, img.ctypes.data_as(C.c_void_p), w, h,
                                         C.c_size_t(w * 3), None, C.c_size_t(0)))
        return img


ROTATE_90_CLOCKWISE, ROTATE_180 = 0, 1


def rotate(src, rotate_code):
    """cv2.rotate(src, ROTATE_90_CLOCKWISE | ROTATE_180) for 8UC1 / 8UC3 (image_stitching.cpp:1093-1103)."""
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape[:2]
    ch = 1 if src.ndim == 2 else src.shape[2]
    shape = (w, h) if rotate_code == ROTATE_90_CLOCKWISE else (h, w)
    dst = np.empty(shape if src.ndim == 2 else shape + (ch,), np.uint8)
    _chk(lib().isb_rotate(src.ctypes.data_as(C.c_void_p), w, h, ch, C.c_size_t(w * ch), int(rotate_code),
                          dst.ctypes.data_as(C.c_void_p), C.c_size_t(shape[1] * ch)))
    return dst


def resize_linear_exact(src, dsize=None, fx=0.0, fy=0.0):
    """cv2.resize(src, dsize, fx=fx, fy=fy, interpolation=INTER_LINEAR_EXACT) for 8UC1 / 8UC3."""
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape[:2]
    ch = 1 if src.ndim == 2 else src.shape[2]
    if dsize is None:
        dw, dh = int(np.rint(w * fx)), int(np.rint(h * fy))  # cvRound
    else:
        dw, dh, fx, fy = int(dsize[0]), int(dsize[1]), 0.0, 0.0
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, ch), np.uint8)
    _chk(lib().isb_resize_linear_exact(src.ctypes.data_as(C.c_void_p), w, h, ch, w * ch, dst.ctypes.data_as(C.c_void_p),
                                       dw, dh, dw * ch, float(fx), float(fy)))
    return dst


def seam_mask_apply(seam_mask, mask_warped):
    """dilate(seam) -> resize(INTER_LINEAR_EXACT, mask_warped.size) -> & mask_warped (image_stitching.cpp:1169-1171)."""
    s = np.ascontiguousarray(seam_mask, np.uint8)
    m = np.ascontiguousarray(mask_warped, np.uint8).copy()
    _chk(lib().isb_seam_mask_apply(s.ctypes.data_as(C.c_void_p), s.shape[1], s.shape[0], C.c_size_t(s.shape[1]),
                                   m.ctypes.data_as(C.c_void_p), m.shape[1], m.shape[0], C.c_size_t(m.shape[1])))
    return m


# ---------------------------------------------------------------------------------------------------
# cv::detail::MultiBandBlender
# ---------------------------------------------------------------------------------------------------
def resultRoi(corners, sizes):
    c = np.ascontiguousarray(corners, np.int32).reshape(-1, 2)
    s = np.ascontiguousarray(sizes, np.int32).reshape(-1, 2)
    r = (C.c_int * 4)()
    _chk(lib().isb_result_roi(c.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), len(c), r))
    return tuple(r)


def num_bands_for(dst_w, dst_h, blend_strength=5.0):
    return int(lib().isb_num_bands_for(int(dst_w), int(dst_h), float(blend_strength)))


class MultiBandBlender:
    def __init__(self, try_gpu=0, num_bands=5):
        self._h = C.c_void_p(lib().isb_blender_create(int(num_bands)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().isb_blender_destroy(self._h)
            self._h = None

    def setNumBands(self, nb):
        _chk(lib().isb_blender_set_num_bands(self._h, int(nb)))

    def numBands(self):
        return int(lib().isb_blender_num_bands(self._h))

    def actualNumBands(self):
        return int(lib().isb_blender_actual_num_bands(self._h))

    def prepare(self, *args):
        if len(args) == 1:
            r = (C.c_int * 4)(*[int(v) for v in args[0]])
            _chk(lib().isb_blender_prepare_roi(self._h, r))
        else:
            c = np.ascontiguousarray(args[0], np.int32).reshape(-1, 2)
            s = np.ascontiguousarray(args[1], np.int32).reshape(-1, 2)
            _chk(lib().isb_blender_prepare(self._h, c.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), len(c)))

    def rois(self):
        a, b = (C.c_int * 4)(), (C.c_int * 4)()
        _chk(lib().isb_blender_get_rois(self._h, a, b))
        return tuple(a), tuple(b)

    def tile_rect(self, w, h, tl):
        r = (C.c_int * 4)()
        _chk(lib().isb_blender_tile_rect(self._h, int(w), int(h), int(tl[0]), int(tl[1]), r))
        return tuple(r)

    def feed(self, img, mask, tl):
        if not _is_torch(img) and np.asarray(img).dtype != np.int16:
            raise IsbError(-215, "Assertion failed: img.type() == CV_16SC3")
        if not _is_torch(mask) and np.asarray(mask).dtype != np.uint8:
            raise IsbError(-215, "Assertion failed: mask.type() == CV_8U")
        ip, _i = _ptr(img)
        mp, _m = _ptr(mask)
        h, w = _shape(mask)[:2]
        _chk(lib().isb_blender_feed(self._h, C.c_void_p(ip), C.c_size_t(w * 6), C.c_void_p(mp), C.c_size_t(w), w, h,
                                    int(tl[0]), int(tl[1])))

    def blend(self, dst=None, dst_mask=None):
        _, rf = self.rois()
        out = np.empty((rf[3], rf[2], 3), np.int16)
        m = np.empty((rf[3], rf[2]), np.uint8)
        _chk(lib().isb_blender_blend(self._h, out.ctypes.data_as(C.c_void_p), C.c_size_t(rf[2] * 6),
                                     m.ctypes.data_as(C.c_void_p), C.c_size_t(rf[2])))
        return out, m


BLENDER_NO, BLENDER_FEATHER, BLENDER_MULTI_BAND = 0, 1, 2  # cv::detail::Blender::{NO, FEATHER, MULTI_BAND}


def createWeightMap(mask, sharpness):
    """cv2.detail.createWeightMap(mask, sharpness, None): min(1, sharpness * L1 distance to the nearest zero pixel)."""
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape
    out = np.empty((h, w), np.float32)
    _chk(lib().isb_create_weight_map(mask.ctypes.data_as(C.c_void_p), w, w, h, float(sharpness),
                                     out.ctypes.data_as(C.c_void_p), w * 4))
    return out


class _SimpleBlender:
    """cv2.detail.Blender_createDefault(Blender_NO) / cv2.detail_FeatherBlender (image_stitching.cpp:1175-1191)."""

    def __init__(self, btype, sharpness):
        h = lib().isb_simple_blender_create(int(btype), float(sharpness))
        if not h:
            raise IsbError(-5, lib().isb_last_error().decode())
        self._h = C.c_void_p(h)
        self._roi = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().isb_simple_blender_destroy(self._h)
            self._h = None

    def prepare(self, *args):
        if len(args) == 1:
            self._roi = tuple(int(v) for v in args[0])
        else:
            self._roi = tuple(resultRoi(args[0], args[1]))
        _chk(lib().isb_simple_blender_prepare_roi(self._h, (C.c_int * 4)(*self._roi)))

    def feed(self, img, mask, tl):
        if not _is_torch(img) and np.asarray(img).dtype != np.int16:
            raise IsbError(-215, "Assertion failed: img.type() == CV_16SC3")
        if not _is_torch(mask) and np.asarray(mask).dtype != np.uint8:
            raise IsbError(-215, "Assertion failed: mask.type() == CV_8U")
        ip, _i = _ptr(img)
        mp, _m = _ptr(mask)
        h, w = _shape(mask)[:2]
        _chk(lib().isb_simple_blender_feed(self._h, C.c_void_p(ip), C.c_size_t(w * 6), C.c_void_p(mp), C.c_size_t(w), w, h,
                                           int(tl[0]), int(tl[1])))

    def blend(self, dst=None, dst_mask=None):
        if self._roi is None:
            raise IsbError(-215, "Assertion failed: prepare() must be called before blend()")
        w, h = self._roi[2], self._roi[3]
        out = np.empty((h, w, 3), np.int16)
        m = np.empty((h, w), np.uint8)
        _chk(lib().isb_simple_blender_blend(self._h, out.ctypes.data_as(C.c_void_p), C.c_size_t(w * 6),
                                            m.ctypes.data_as(C.c_void_p), C.c_size_t(w)))
        return out, m


class FeatherBlender(_SimpleBlender):
    def __init__(self, sharpness=0.02):
        super().__init__(BLENDER_FEATHER, sharpness)

    def setSharpness(self, val):
        _chk(lib().isb_simple_blender_set_sharpness(self._h, float(val)))

    def sharpness(self):
        return float(lib().isb_simple_blender_sharpness(self._h))


def Blender_createDefault(btype, try_gpu=False):
    """cv2.detail.Blender_createDefault: NO -> plain Blender, FEATHER -> FeatherBlender(), MULTI_BAND -> MultiBandBlender()."""
    if btype == BLENDER_NO:
        return _SimpleBlender(BLENDER_NO, 0.02)
    if btype == BLENDER_FEATHER:
        return FeatherBlender()
    if btype == BLENDER_MULTI_BAND:
        return MultiBandBlender()
    raise IsbError(-5, "unknown blender type")


TIMELAPSER_AS_IS, TIMELAPSER_CROP = 0, 1  # cv::detail::Timelapser::{AS_IS, CROP}


class Timelapser:
    """cv2.detail.Timelapser_createDefault(type): initialize / process / getDst (image_stitching.cpp:1194-1215)."""

    def __init__(self, ttype=TIMELAPSER_AS_IS):
        h = lib().isb_timelapser_create(int(ttype))
        if not h:
            raise IsbError(-5, lib().isb_last_error().decode())
        self._h = C.c_void_p(h)
        self.dst_roi = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().isb_timelapser_destroy(self._h)
            self._h = None

    def initialize(self, corners, sizes):
        c = np.ascontiguousarray(corners, np.int32).reshape(-1, 2)
        s = np.ascontiguousarray(sizes, np.int32).reshape(-1, 2)
        roi = (C.c_int * 4)()
        _chk(lib().isb_timelapser_initialize(self._h, c.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), len(c), roi))
        self.dst_roi = tuple(roi)

    def process(self, img, mask, tl):
        if not _is_torch(img) and np.asarray(img).dtype != np.int16:
            raise IsbError(-215, "Assertion failed: img.type() == CV_16SC3")
        ip, _i = _ptr(img)
        h, w = _shape(img)[:2]
        _chk(lib().isb_timelapser_process(self._h, C.c_void_p(ip), C.c_size_t(w * 6), w, h, int(tl[0]), int(tl[1])))

    def getDst(self):
        w, h = self.dst_roi[2], self.dst_roi[3]
        out = np.zeros((h, w, 3), np.int16)
        if w > 0 and h > 0:
            _chk(lib().isb_timelapser_get_dst(self._h, out.ctypes.data_as(C.c_void_p), C.c_size_t(w * 6)))
        return out


def Timelapser_createDefault(ttype):
    return Timelapser(ttype)


# ---------------------------------------------------------------------------------------------------
# fused loop
# ---------------------------------------------------------------------------------------------------
def crop_rect(mask, with_points=False):
    """Rectangle (x, y, w, h) the reference's crop() (cropper.cpp:116-209) narrows the panorama to, from its 8UC1 mask
    (numpy or torch.cuda, 2-D)."""
    p, keep = _ptr(mask)
    h, w = _shape(mask)[:2]
    r = (C.c_int * 4)()
    n = C.c_int(0)
    _chk(lib().isb_crop_rect(C.c_void_p(p), int(w), int(h), C.c_size_t(int(w)), r, C.byref(n)))
    return (tuple(r), n.value) if with_points else tuple(r)


def crop(image):
    """crop(source) of the reference: returns (cropped view, rect).  image: HxWx3 uint8 or int16 numpy array."""
    a = np.ascontiguousarray(image)
    assert a.ndim == 3 and a.shape[2] == 3 and a.dtype in (np.uint8, np.int16)
    r = (C.c_int * 4)()
    _chk(lib().isb_crop_rect_image(a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], C.c_size_t(a.strides[0]),
                                   int(a.dtype == np.int16), r, None))
    x, y, w, h = tuple(r)
    out = np.clip(a, 0, 255).astype(np.uint8) if a.dtype == np.int16 else a  # crop() converts the source to CV_8U
    return out[y:y + h, x:x + w], (x, y, w, h)


def imencode_jpg(image, quality=95):
    """cv2.imencode('.jpg', image) / the file cv::imwrite("result.jpg", result) writes (image_stitching.cpp:1228): bytes.
    image: HxWx3 uint8 or int16 (numpy, BGR), or a torch.cuda uint8 / int16 tensor."""
    if hasattr(image, "data_ptr"):
        assert image.is_contiguous() and image.dim() == 3 and image.shape[2] == 3
        h, w = int(image.shape[0]), int(image.shape[1])
        is16 = int(image.element_size() == 2)
        ptr, pitch = C.c_void_p(image.data_ptr()), w * 3 * image.element_size()
        keep = image
    else:
        keep = np.ascontiguousarray(image)
        assert keep.ndim == 3 and keep.shape[2] == 3 and keep.dtype in (np.uint8, np.int16)
        h, w = keep.shape[:2]
        is16 = int(keep.dtype == np.int16)
        ptr, pitch = keep.ctypes.data_as(C.c_void_p), keep.strides[0]
    cap = 1024 + w * h * 3 // 2  # room for ordinary images; the call reports the size it needs when this is too small
    n = C.c_size_t(0)
    for _ in range(2):
        buf = np.empty(cap, np.uint8)
        rc = lib().isb_jpeg_encode(ptr, w, h, C.c_size_t(pitch), is16, int(quality), buf.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n))
        if rc == 0:
            return buf[:n.value].tobytes()
        if n.value <= cap:
            _chk(rc)
        cap = int(n.value)
    _chk(rc)


def cameras_from_KR(Ks, Rs):
    """isb_camera list from float32 K = [[f,0,cx],[0,f*a,cy],[0,0,1]] and R."""
    cams = []
    for K, R in zip(Ks, Rs):
        K = np.asarray(K, np.float64)
        cams.append(Camera.make(K[0, 0], K[0, 2], K[1, 2], R, aspect=K[1, 1] / K[0, 0]))
    return cams


GATHER_PEER_STORES, GATHER_COPY_ENGINE, GATHER_LOCAL = 0, 1, 2


class Composer:
    """The whole compositing loop on the GPU (isb_composer_*)."""

    def __init__(self, warp="spherical", scale=1.0, num_bands=5, strip_index=0, strip_count=1, cache_plan=True,
                 async_mode=False, gather_copy=False, gather_mode=None, pipeline_depth=1, blend_type=None, blend_strength=5.0, ingest_rotate=None,
                 compose_scale=0.0):
        self.cfg = Config()
        self.cfg.warp_kind = _KIND[warp]
        self.cfg.warped_image_scale = float(scale)
        self.cfg.num_bands = int(num_bands)
        self.cfg.strip_index, self.cfg.strip_count = int(strip_index), int(strip_count)
        self.cfg.cache_plan = int(bool(cache_plan))
        self.cfg.async_mode = int(bool(async_mode))
        # GATHER_PEER_STORES (0) / GATHER_COPY_ENGINE (1) / GATHER_LOCAL (2); gather_copy=True is shorthand for 1
        self.cfg.gather_mode = int(gather_mode) if gather_mode is not None else (GATHER_COPY_ENGINE if gather_copy else GATHER_PEER_STORES)
        self.cfg.pipeline_depth = int(pipeline_depth)
        # ingest pre-steps inside the composer: images passed to run() are the decoded frames (rotate code as for isb.rotate)
        self.cfg.ingest_rotate = 0 if ingest_rotate is None else 1 + int(ingest_rotate)
        self.cfg.compose_scale = float(compose_scale)
        if blend_type is not None:  # the reference's blender set-up (image_stitching.cpp:1173-1193) instead of explicit num_bands
            self.cfg.use_blend_rule = 1
            self.cfg.blend_type = {"no": BLENDER_NO, "feather": BLENDER_FEATHER, "multiband": BLENDER_MULTI_BAND}.get(blend_type, blend_type)
            self.cfg.blend_strength = float(blend_strength)
        self._h = C.c_void_p(lib().isb_composer_create(C.byref(self.cfg)))
        self.n = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().isb_composer_destroy(self._h)
            self._h = None

    def plan(self, cams, src_sizes_wh):
        n = len(cams)
        arr = (Camera * n)(*cams)
        sz = np.ascontiguousarray(src_sizes_wh, np.int32).reshape(n, 2)
        corners, sizes = np.zeros((n, 2), np.int32), np.zeros((n, 2), np.int32)
        roi = (C.c_int * 4)()
        _chk(lib().isb_composer_plan(self._h, arr, sz.ctypes.data_as(C.c_void_p), n, corners.ctypes.data_as(C.c_void_p),
                                     sizes.ctypes.data_as(C.c_void_p), roi))
        self.n = n
        self.corners = [tuple(int(v) for v in c) for c in corners]
        self.sizes = [tuple(int(v) for v in s) for s in sizes]
        self.dst_roi = tuple(roi)
        return self.corners, self.sizes, self.dst_roi

    def run(self, images, gains=None, seam_masks=None, out=None, out_mask=None, out16=None, want16=False, out_pitch=None,
            mask_pitch=None):
        """images: list of HxWx3 uint8 (numpy or torch.cuda).  Outputs are allocated (numpy) unless given; out_pitch /
        mask_pitch (bytes) describe caller buffers whose rows are wider than the panorama."""
        n = self.n
        keep = []
        ia = (_Image * n)()
        for i, im in enumerate(images):
            p, k = _ptr(im)
            keep.append(k)
            h, w = _shape(im)[:2]
            ia[i] = _Image(p, w, h, w * 3)
        ga = None
        if gains is not None:
            ga = (_Gain * n)()
            for i, g in enumerate(gains):
                if g is None:
                    ga[i] = _Gain(None, 0, 0)
                    continue
                if not _is_torch(g):
                    g = np.ascontiguousarray(g, np.float32)
                p, k = _ptr(g)
                keep.append(k)
                ga[i] = _Gain(p, _shape(g)[1], _shape(g)[0])
        sa = None
        if seam_masks is not None:
            sa = (_Mask * n)()
            for i, m in enumerate(seam_masks):
                if m is None:
                    sa[i] = _Mask(None, 0, 0, 0)
                    continue
                p, k = _ptr(m)
                keep.append(k)
                sa[i] = _Mask(p, _shape(m)[1], _shape(m)[0], _shape(m)[1])
        x, y, w, h = self.dst_roi
        if out is None:
            out = np.zeros((h, w, 3), np.uint8)
        if out_mask is None:
            out_mask = np.zeros((h, w), np.uint8)
        if out16 is None and want16:
            out16 = np.zeros((h, w, 3), np.int16)
        pano = _Pano()
        pano.data, k1 = _ptr(out)
        pano.pitch = w * 3 if out_pitch is None else int(out_pitch)
        pano.mask, k2 = _ptr(out_mask)
        pano.mask_pitch = w if mask_pitch is None else int(mask_pitch)
        if out16 is not None:
            pano.data16, k3 = _ptr(out16)
            pano.pitch16 = w * 6
        _chk(lib().isb_composer_run(self._h, ia, ga, sa, n, C.byref(pano)))
        self.strip_rows = (pano.strip_y0, pano.strip_y1)
        return dict(result8=out, mask=out_mask, result16=out16, dst_roi=tuple(pano.roi_xywh),
                    strip_rows=self.strip_rows, corners=self.corners, sizes=self.sizes)

    def sync(self):
        _chk(lib().isb_composer_sync(self._h))

    def join(self):
        """Stream-ordered wait (no host sync) for the copy-engine gather of all previous runs."""
        _chk(lib().isb_composer_join(self._h))

    def timings(self):
        ms = (C.c_float * 8)()
        n = lib().isb_composer_last_timings(self._h, ms, 8)
        if n < 0:
            _chk(n)
        return {lib().isb_composer_stage_name(i).decode(): float(ms[i]) for i in range(n)}

    def planned_rows(self):
        """Rows (y0, y1) of the panorama this composer's strip covers (valid after plan())."""
        a, b = C.c_int(0), C.c_int(0)
        _chk(lib().isb_composer_strip_rows(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def source_band(self, index):
        """Rows [lo, hi] of source image `index` this strip reads (strip-sharded plans; the whole image otherwise)."""
        lo, hi = C.c_int(0), C.c_int(0)
        _chk(lib().isb_composer_source_band(self._h, int(index), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def last_h2d_bytes(self):
        return int(lib().isb_composer_last_h2d_bytes(self._h))

    def byte_model(self):
        S, M, A, B = C.c_double(0), C.c_double(0), C.c_double(0), C.c_double(0)
        _chk(lib().isb_composer_byte_model(self._h, C.byref(S), C.byref(M), C.byref(A), C.byref(B)))
        return dict(S=S.value, M=M.value, Ap=A.value, B_alg=B.value)


def compose(images, Ks, Rs, scale, warp, num_bands, gains=None, seam_masks=None, want16=True, **kw):
    """One-shot fused loop with the same result dict as the oracle's compose()."""
    c = Composer(warp, scale, num_bands, **kw)
    c.plan(cameras_from_KR(Ks, Rs), [(_shape(im)[1], _shape(im)[0]) for im in images])
    return c.run(images, gains, seam_masks, want16=want16)


def compose_with_blender(images, Ks, Rs, scale, warp, blender, gains=None, seam_masks=None):
    """The compositing loop (image_stitching.cpp:1086-1229) call by call through the C ABI mirrors, for any blender with the
    prepare / feed / blend contract (MultiBandBlender, FeatherBlender, Blender_createDefault(BLENDER_NO)): warper.warp of the
    image and of the all-255 mask, compensator.apply, convertTo(16S), dilate + INTER_LINEAR_EXACT resize + AND, feed, blend."""
    warper = RotationWarper(warp, scale)
    corners, sizes = [], []
    for im, K, R in zip(images, Ks, Rs):
        h, w = _shape(im)[:2]
        x, y, rw, rh = warper.warpRoi((w, h), K, R)
        corners.append((x, y))
        sizes.append((rw, rh))
    comp = None
    if gains is not None:
        comp = BlocksGainCompensator(64, 64, 1)
        comp.setMatGains(gains)
    blender.prepare(corners, sizes)
    for i, (im, K, R) in enumerate(zip(images, Ks, Rs)):
        _, img_warped = warper.warp(im, K, R, INTER_LINEAR, BORDER_REFLECT)
        _, mask_warped = warper.warp(np.full(_shape(im)[:2], 255, np.uint8), K, R, INTER_NEAREST, BORDER_CONSTANT)
        if comp is not None:
            img_warped = comp.apply(i, corners[i], img_warped, mask_warped)
        if seam_masks is not None:
            mask_warped = seam_mask_apply(seam_masks[i], mask_warped)
        blender.feed(img_warped.astype(np.int16), mask_warped, corners[i])
    result, result_mask = blender.blend()
    return dict(corners=corners, sizes=sizes, dst_roi=tuple(resultRoi(corners, sizes)), result16=result,
                result8=np.clip(result, 0, 255).astype(np.uint8), mask=result_mask)


def strip_rows(padded_h, final_h, num_bands, index, count):
    a, b = C.c_int(0), C.c_int(0)
    _chk(lib().isb_strip_rows(int(padded_h), int(final_h), int(num_bands), int(index), int(count), C.byref(a), C.byref(b)))
    return a.value, b.value

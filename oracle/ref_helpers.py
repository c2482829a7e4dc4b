"""ctypes binding of oracle/_ref/libisb_ref.so - the reference's OWN helper sources (quaternion.h, euler.h,
serializer.cpp, cropper.cpp) compiled from /root/reference by `make -C oracle _ref` (TEST INFRASTRUCTURE ONLY).

The library is built in the build container (where /root/reference exists) and travels to the GPU box as a
git-ignored artefact; `available()` tells the tests whether it is there."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libisb_ref.so")
REF_DIR = "/root/reference/image_stitching"
EULER = {"XYZ": 0, "YXZ": 1, "ZXY": 2, "ZYX": 3, "YZX": 4, "XZY": 5}
_LIB = None
_KEEP = []  # ctypes callbacks must outlive their registration
_BUFS = []  # arrays handed to the C side by the last few findContours calls


def build(force: bool = False):
    """(Re)build when the reference tree is present; otherwise the prebuilt library (if any) is used as is."""
    if os.path.isdir(REF_DIR):
        deps = [os.path.join(_HERE, "ref_glue.cpp"), os.path.join(_HERE, "Makefile")]
        for root, _, files in os.walk(os.path.join(_HERE, "cvshim")):
            deps += [os.path.join(root, f) for f in files]
        stale = not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps)
        if force or stale:
            subprocess.check_call(["make", "-C", _HERE, "-B", "_ref"], stdout=subprocess.DEVNULL)
    return SO if os.path.exists(SO) else None


def available() -> bool:
    return build() is not None


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libisb_ref.so is missing and /root/reference is not here to build it")
        _LIB = C.CDLL(so)
    return _LIB


def _d(a, n):
    a = np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1))
    assert a.size == n
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def quat_from_rotation_matrix(R):
    R, q = _d(R, 9), np.zeros(4)
    lib().ref_quat_from_rotation_matrix(_p(R), _p(q))
    return q


def quat_to_rotation_matrix(q):
    q, R = _d(q, 4), np.zeros(9)
    lib().ref_quat_to_rotation_matrix(_p(q), _p(R))
    return R.reshape(3, 3)


def quat_from_euler(e, order):
    e, q = _d(e, 3), np.zeros(4)
    lib().ref_quat_from_euler(_p(e), EULER[order], _p(q))
    return q


def quat_from_axis_angle(axis, angle):
    a, q = _d(axis, 3), np.zeros(4)
    lib().ref_quat_from_axis_angle(_p(a), C.c_double(angle), _p(q))
    return q


def quat_multiply(a, b):
    a, b, q = _d(a, 4), _d(b, 4), np.zeros(4)
    lib().ref_quat_multiply(_p(a), _p(b), _p(q))
    return q


def quat_slerp(a, b, t):
    a, b, q = _d(a, 4), _d(b, 4), np.zeros(4)
    lib().ref_quat_slerp(_p(a), _p(b), C.c_double(t), _p(q))
    return q


def pose_from_cam_transform(R, is_portrait):
    R, o = _d(R, 9), np.zeros(9)
    lib().ref_pose_from_cam_transform(_p(R), int(bool(is_portrait)), _p(o))
    return o.reshape(3, 3)


def rotationMatrixToEulerAngles(R, order):
    R, e = _d(R, 9), np.zeros(3)
    lib().ref_rotation_matrix_to_euler(_p(R), EULER[order], _p(e))
    return e


def eulerAnglesToRotationMatrix(e, order):
    e, R = _d(e, 3), np.zeros(9)
    lib().ref_euler_to_rotation_matrix(_p(e), EULER[order], _p(R))
    return R.reshape(3, 3)


def parseMatrixStr(s):
    out = np.zeros(4096)
    side = lib().ref_parse_matrix_str(s.encode(), _p(out), out.size)
    return out[: side * side].reshape(side, side).copy()


def serializeMatrix(m):
    m = np.asarray(m)
    is32 = m.dtype == np.float32
    md = np.ascontiguousarray(m, np.float64)
    rows, cols = (md.shape + (1,))[:2] if md.ndim == 1 else md.shape
    buf = C.create_string_buffer(64 * md.size + 16)
    n = lib().ref_serialize_matrix(_p(md), int(rows), int(cols), int(is32), buf, len(buf))
    assert n >= 0
    return buf.value.decode()


def deserializeMatrix(s):
    out = np.zeros(4096, np.float32)
    r, c = C.c_int(0), C.c_int(0)
    lib().ref_deserialize_matrix(s.encode(), _p(out), out.size, C.byref(r), C.byref(c))
    return out[: r.value * c.value].reshape(r.value, c.value).copy()


class _Chdir:
    def __init__(self, d):
        self.d = d

    def __enter__(self):
        self.old = os.getcwd()
        os.chdir(self.d)

    def __exit__(self, *a):
        os.chdir(self.old)


def save_cams(cam_array, n, directory):
    """serializeCameraParams writes ./cams.data - run it inside `directory`.  cam_array: ctypes array of isb.Camera."""
    with _Chdir(directory):
        lib().ref_save_cams(cam_array, int(n))


def load_cams(cam_array, cap, directory):
    with _Chdir(directory):
        return int(lib().ref_load_cams(cam_array, int(cap)))


def save_indices(idx, directory):
    a = np.ascontiguousarray(idx, np.int32)
    with _Chdir(directory):
        lib().ref_save_indices(_p(a), a.size)


def load_indices(directory, cap=4096):
    a = np.zeros(cap, np.int32)
    with _Chdir(directory):
        n = int(lib().ref_load_indices(_p(a), cap))
    return [int(v) for v in a[:n]]


# ---- cropper.cpp: crop() with cv2 answering findContours / drawContours ----------------------------------------------
_FIND = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)),
                    C.POINTER(C.c_int))
_DRAW = C.CFUNCTYPE(None, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int)


def install_cv2_contours():
    """findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) and drawContours(filled) of the real OpenCV behind the shim."""
    import cv2

    def find(mask, w, h, xy_out, lens_out, n_out):
        m = np.ctypeslib.as_array(mask, shape=(h, w)).copy()
        contours, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
        xy = np.ascontiguousarray(np.concatenate([c.reshape(-1, 2) for c in contours]) if contours else np.zeros((0, 2)), np.int32)
        lens = np.ascontiguousarray([len(c) for c in contours], np.int32)
        _BUFS.append((xy, lens))
        del _BUFS[:-4]
        xy_out[0] = xy.ctypes.data_as(C.POINTER(C.c_int))
        lens_out[0] = lens.ctypes.data_as(C.POINTER(C.c_int))
        n_out[0] = len(contours)
        return 0

    def draw(img, w, h, xy, npts):
        pts = np.ctypeslib.as_array(xy, shape=(npts, 2)).astype(np.int32).reshape(-1, 1, 2)
        m = np.zeros((h, w), np.uint8)
        cv2.drawContours(m, [pts], 0, 255, -1, 8)
        np.ctypeslib.as_array(img, shape=(h, w))[:] = m

    f, d = _FIND(find), _DRAW(draw)
    _KEEP.append((f, d))
    lib().ref_set_contour_callbacks(f, d)


def crop_rect(img):
    """Rectangle (x, y, w, h) the reference's crop(source) narrows `source` to.  img: 8UC3 or 16SC3."""
    a = np.ascontiguousarray(img)
    assert a.ndim == 3 and a.shape[2] == 3 and a.dtype in (np.uint8, np.int16)
    r = np.zeros(4, np.int32)
    lib().ref_crop(_p(a), a.shape[1], a.shape[0], int(a.dtype == np.int16), _p(r))
    return tuple(int(v) for v in r)


def check_interior_exterior(mask, rect):
    m = np.ascontiguousarray(mask, np.uint8)
    r = np.ascontiguousarray(rect, np.int32)
    o = np.zeros(4, np.int32)
    ok = lib().ref_check_interior_exterior(_p(m), m.shape[1], m.shape[0], _p(r), _p(o))
    return bool(ok), tuple(int(v) for v in o)

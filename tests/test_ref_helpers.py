"""a11 pinned against the reference's OWN code: image_stitching/quaternion.h, euler.h and serializer.cpp compiled
unmodified from /root/reference into oracle/_ref/libisb_ref.so (oracle/Makefile target _ref, cv core types from
oracle/cvshim).  Every product host helper must agree with it BIT FOR BIT on random inputs, including the gimbal
branches of euler.h:4-133, all four branches of quaternion.h:260-322 and the EXIF pose fix-up image_stitching.cpp:485-517."""
import math

import numpy as np
import pytest

import image_stitching_b200 as isb
from image_stitching_b200 import synth
from oracle import ref_helpers as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libisb_ref.so not built (no /root/reference here)")

ORDERS = ["XYZ", "YXZ", "ZXY", "ZYX", "YZX", "XZY"]


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def same(a, b):
    return np.array_equal(bits(a), bits(b))


def random_rotation(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    return isb.quat_to_rotation_matrix(q)


def test_euler_matches_reference_bitwise():
    rng = np.random.default_rng(11)
    for order in ORDERS:
        for k in range(200):
            e = rng.uniform(-math.pi, math.pi, 3)
            if k % 10 == 0:  # gimbal lock: middle angle at +-pi/2 (euler.h's `< 0.9999999` guard takes the else branch)
                e[{"XYZ": 1, "YXZ": 0, "ZXY": 0, "ZYX": 1, "YZX": 2, "XZY": 2}[order]] = (math.pi / 2) * (1 if k % 20 else -1)
            R_ref = ref.eulerAnglesToRotationMatrix(e, order)
            R = isb.eulerAnglesToRotationMatrix(e, order)
            assert same(R, R_ref), (order, e)
            assert same(isb.rotationMatrixToEulerAngles(R, order), ref.rotationMatrixToEulerAngles(R_ref, order)), (order, e)
        for _ in range(50):  # arbitrary (not exactly orthonormal) matrices, entries beyond +-1: the clamp
            M = rng.uniform(-1.3, 1.3, (3, 3))
            assert same(isb.rotationMatrixToEulerAngles(M, order), ref.rotationMatrixToEulerAngles(M, order))
    # the rig formula of synth.py (YXZ, row/column convention of euler.h:289-297) is the reference's
    assert same(synth.euler_yxz_to_R(0.3, -0.8, 0.1), ref.eulerAnglesToRotationMatrix([0.3, -0.8, 0.1], "YXZ"))


def test_quaternion_matches_reference_bitwise():
    rng = np.random.default_rng(12)
    seen = set()
    for k in range(400):
        R = random_rotation(rng) if k % 3 else rng.uniform(-1, 1, (3, 3))
        tr = R[0, 0] + R[1, 1] + R[2, 2]
        seen.add(0 if tr > 0 else 1 if (R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]) else 2 if R[1, 1] > R[2, 2] else 3)
        q_ref = ref.quat_from_rotation_matrix(R)
        assert same(isb.quat_from_rotation_matrix(R), q_ref)
        assert same(isb.quat_to_rotation_matrix(q_ref), ref.quat_to_rotation_matrix(q_ref))
        for portrait in (False, True):
            assert same(isb.pose_from_cam_transform(R, portrait), ref.pose_from_cam_transform(R, portrait))
    assert seen == {0, 1, 2, 3}, "all four branches of setFromRotationMatrix must be exercised"
    for order in ORDERS:
        for _ in range(100):
            e = rng.uniform(-4, 4, 3)
            assert same(isb.quat_from_euler(e, order), ref.quat_from_euler(e, order))
    for _ in range(200):
        a, b = rng.standard_normal(4), rng.standard_normal(4)
        assert same(isb.quat_multiply(a, b), ref.quat_multiply(a, b))
        axis = rng.standard_normal(3)
        ang = rng.uniform(-7, 7)
        assert same(isb.quat_from_axis_angle(axis, ang), ref.quat_from_axis_angle(axis, ang))
    for k in range(300):
        a, b = rng.standard_normal(4), rng.standard_normal(4)
        a /= np.linalg.norm(a)
        b /= np.linalg.norm(b)
        if k % 5 == 0:
            b = a.copy()                     # cosHalfTheta >= 1
        elif k % 5 == 1:
            b = -a + 1e-9 * rng.standard_normal(4)   # negative dot product, nearly parallel (linear branch)
        elif k % 5 == 2:
            b = a + 1e-9 * rng.standard_normal(4)    # sqrSinHalfTheta <= eps
        t = [0.0, 1.0, 0.5, rng.uniform(0, 1), rng.uniform(-0.5, 1.5)][k % 5 if k % 7 else 0]
        assert same(isb.quat_slerp(a, b, t), ref.quat_slerp(a, b, t)), (a, b, t)


def test_serializer_matches_reference(tmp_path):
    rng = np.random.default_rng(13)
    for _ in range(100):
        r, c = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        m = (rng.standard_normal((r, c)) * 10.0 ** rng.integers(-8, 9, (r, c)))
        for dt in (np.float32, np.float64):
            s_ref = ref.serializeMatrix(m.astype(dt))
            assert isb.serializeMatrix(m.astype(dt)) == s_ref
            back_ref = ref.deserializeMatrix(s_ref)
            back = isb.deserializeMatrix(s_ref)
            assert back.shape == back_ref.shape and np.array_equal(back.view(np.uint32), back_ref.view(np.uint32))
    for side in (1, 2, 3, 4):
        vals = rng.standard_normal(side * side) * 100
        s = "[" + ",".join(repr(float(v)) for v in vals) + "]"
        assert same(isb.parseMatrixStr(s), ref.parseMatrixStr(s))
    # the EXIF payload form: 16 values -> 4x4 (serializer.cpp:22-36)
    s = "[1,0,0,0,0,1,0,0,0,0,1,0,0.5,0.25,2,1]"
    assert same(isb.parseMatrixStr(s), ref.parseMatrixStr(s))

    # cams.data: product writer -> reference reader, reference writer -> product reader, and byte-identical files
    rig = synth.make_rig("cfg3", scale_div=8)
    cams = isb.cameras_from_KR(rig.Ks, rig.Rs)
    for i, c in enumerate(cams):
        c.t[:] = [0.5 * i, -1.25, 3.0 + i / 7.0]
        c.aspect = 1.0 + 0.01 * i
    arr = (isb.Camera * len(cams))(*cams)
    d_ref, d_our = tmp_path / "ref", tmp_path / "our"
    d_ref.mkdir()
    d_our.mkdir()
    ref.save_cams(arr, len(cams), str(d_ref))
    isb.serializeCameraParams(cams, str(d_our / "cams.data"))
    assert (d_ref / "cams.data").read_bytes() == (d_our / "cams.data").read_bytes()
    back_our = isb.deserializeCameraParams(str(d_ref / "cams.data"))
    back_ref = (isb.Camera * len(cams))()
    assert ref.load_cams(back_ref, len(cams), str(d_our)) == len(cams) == len(back_our)
    for a, b in zip(back_our, back_ref):
        assert (a.focal, a.aspect, a.ppx, a.ppy) == (b.focal, b.aspect, b.ppx, b.ppy)
        assert list(a.R) == list(b.R) and list(a.t) == list(b.t)
    # indices.data
    idx = [3, 0, 7, 12345, -2]
    ref.save_indices(idx, str(d_ref))
    isb.serializeIndices(idx, str(d_our / "indices.data"))
    assert (d_ref / "indices.data").read_bytes() == (d_our / "indices.data").read_bytes()
    (d_ref / "indices.data").write_text("3\n\n0\n7\n\n9\n")
    assert isb.deserializeIndices(str(d_ref / "indices.data")) == ref.load_indices(str(d_ref)) == [3, 0, 7, 9]

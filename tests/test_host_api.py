"""CPU: the C-ABI library loads, exports every symbol include/image_stitching.h declares, and its host-side
logic (integer geometry, pose math, serializer, strip planner) matches the oracle.  No compute calls."""
import os
import re

import numpy as np
import pytest
from conftest import ROOT

import image_stitching_b200 as isb
from image_stitching_b200 import synth
from oracle import oracle as orc


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "image_stitching.h")).read()
    names = sorted(set(re.findall(r"ISB_API[^;(]*?\b(isb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 40
    lib = isb.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_compute_fails_loudly_without_gpu():
    if isb.device_count() > 0:
        pytest.skip("a GPU is present")
    w = isb.RotationWarper("spherical", 100.0)
    K = np.array([[100, 0, 50], [0, 100, 40], [0, 0, 1]], np.float32)
    with pytest.raises(isb.IsbError) as e:
        w.warp(np.zeros((80, 100, 3), np.uint8), K, np.eye(3, dtype=np.float32), isb.INTER_LINEAR, isb.BORDER_REFLECT)
    assert e.value.code == -217 and "no CPU fallback" in str(e.value)
    b = isb.MultiBandBlender(0, 3)
    b.prepare((0, 0, 100, 80))
    with pytest.raises(isb.IsbError):
        b.feed(np.zeros((80, 100, 3), np.int16), np.zeros((80, 100), np.uint8), (0, 0))


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4", "cfg5"])
def test_roi_geometry_full_size(name):
    """a1/a7 of SURVEY.md 8(a): bit-exact integer geometry at BASELINE.json's full sizes."""
    expect = {"cfg2": (-10455, 3785, 20912, 2881), "cfg3": (-23327, 4713, 46655, 13903),
              "cfg4": (-6465, -1111, 10550, 2212), "cfg5": (-42132, 53943, 82488, 32653)}[name]
    rig = synth.make_rig(name)
    w = isb.RotationWarper(rig.warp, rig.scale)
    n = min(rig.n, 40)
    rois = [w.warpRoi((rig.W, rig.H), K, R) for K, R in zip(rig.Ks[:n], rig.Rs[:n])]
    assert rois == [orc.warp_roi(rig.warp, rig.scale, rig.W, rig.H, K, R) for K, R in zip(rig.Ks[:n], rig.Rs[:n])]
    if n == rig.n:
        assert isb.resultRoi([r[:2] for r in rois], [r[2:] for r in rois]) == expect
    # composer geometry pass (host only) agrees
    c = isb.Composer(rig.warp, rig.scale, rig.nb)
    if isb.device_count() == 0:
        with pytest.raises(isb.IsbError):
            c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)


def test_warp_point_matches_oracle():
    rig = synth.make_rig("cfg2", 4)
    for kind in ("spherical", "cylindrical"):
        w = isb.RotationWarper(kind, rig.scale)
        for K, R in zip(rig.Ks[:3], rig.Rs[:3]):
            for pt in [(0.0, 0.0), (123.5, 77.25), (rig.W - 1.0, rig.H - 1.0)]:
                assert w.warpPoint(pt, K, R) == orc.map_forward(kind, rig.scale, K, R, *pt)
            for uv in [(-100.0, 900.0), (250.5, 1200.0)]:
                assert w.warpPointBackward(uv, K, R) == orc.map_backward(kind, rig.scale, K, R, *uv)


def test_blender_geometry_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(1, 6))
        corners = [(int(rng.integers(-3000, 3000)), int(rng.integers(-500, 500))) for _ in range(n)]
        sizes = [(int(rng.integers(1, 2500)), int(rng.integers(1, 900))) for _ in range(n)]
        nb = int(rng.integers(0, 13))
        roi = isb.resultRoi(corners, sizes)
        assert roi == orc.result_roi(corners, sizes)
        a, b = isb.MultiBandBlender(0, nb), orc.Blender(nb)
        a.prepare(corners, sizes)
        b.prepare(roi)
        assert a.numBands() == nb  # the requested value, as OpenCV reports it
        assert a.actualNumBands() == b.numBands()
        assert a.rois() == b.rois()
        for c, s in zip(corners, sizes):
            assert a.tile_rect(s[0], s[1], c) == b.tile_rect(s[0], s[1], c)


def test_num_bands_rule():
    # image_stitching.cpp:1177-1183 with blend_strength = 5
    import math
    for (w, h) in [(20912, 2881), (46655, 13903), (300, 200), (10, 10)]:
        bw = np.float32(math.sqrt(np.float32(w * h))) * np.float32(5) / np.float32(100)
        want = -1 if bw < 1 else int(math.ceil(math.log(bw) / math.log(2.0)) - 1.0)
        assert isb.num_bands_for(w, h, 5.0) == want


def test_strip_rows_partition_the_panorama():
    for (ph, fh, nb, n) in [(2912, 2881, 5, 8), (13952, 13903, 7, 8), (32768, 32653, 8, 4), (64, 40, 5, 8), (96, 90, 5, 2)]:
        prev = 0
        for i in range(n):
            y0, y1 = isb.strip_rows(ph, fh, nb, i, n)
            assert y0 == prev and y1 >= y0 and (y0 % (1 << nb) == 0 or y0 == fh)
            prev = y1
        assert prev == fh


def test_euler_and_quaternion_contract():
    rng = np.random.default_rng(2)
    for order in isb.EULER:
        for _ in range(20):
            e = rng.uniform(-1.2, 1.2, 3)
            R = isb.eulerAnglesToRotationMatrix(e, order)
            assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
            assert np.allclose(isb.rotationMatrixToEulerAngles(R, order), e, atol=1e-9)
            q = isb.quat_from_rotation_matrix(R)
            assert np.allclose(isb.quat_to_rotation_matrix(q), R, atol=1e-12)
            # setFromEuler(order) and makeRotationFromEuler(order) describe the same rotation (three.js convention)
            q2 = isb.quat_from_euler(e, order)
            assert np.allclose(isb.quat_to_rotation_matrix(q2), R, atol=1e-12)
    # YXZ is the order the reference prints (image_stitching.cpp:731-741); bit-equal with the rig formula
    assert np.array_equal(isb.eulerAnglesToRotationMatrix([0.3, -0.8, 0.1], "YXZ"), synth.euler_yxz_to_R(0.3, -0.8, 0.1))
    a = isb.quat_from_axis_angle([0, 1, 0], 0.7)
    b = isb.quat_from_euler([0, 0.7, 0], "XYZ")
    assert np.allclose(a, b)
    assert np.allclose(isb.quat_multiply(a, [0, 0, 0, 1]), a)
    assert np.allclose(isb.quat_slerp(a, b, 0.5), a)
    h = isb.quat_slerp([0, 0, 0, 1], isb.quat_from_axis_angle([0, 0, 1], 1.0), 0.5)
    assert np.allclose(h, isb.quat_from_axis_angle([0, 0, 1], 0.5))
    # the EXIF pose fix-up (image_stitching.cpp:485-517)
    R = isb.eulerAnglesToRotationMatrix([0.2, 0.5, -0.3], "YXZ")
    q = isb.quat_from_rotation_matrix(R)
    assert np.allclose(isb.pose_from_cam_transform(R, True), isb.quat_to_rotation_matrix([q[1], q[0], -q[2], q[3]]))
    assert np.allclose(isb.pose_from_cam_transform(R, False), isb.quat_to_rotation_matrix([-q[0], q[1], -q[2], q[3]]))


def test_serializer_formats(tmp_path):
    # writer: ',' between columns, ';' after the last column of every row, 6 significant digits (serializer.cpp:38-67)
    assert isb.serializeMatrix(np.array([[1.5, 2, 3], [4, 5, 6.25]], np.float32)) == "[1.5,2,3;4,5,6.25;]"
    assert isb.serializeMatrix(np.array([[0.1234567891], [2e-7], [12345678.0]], np.float64)) == "[0.123457;2e-07;1.23457e+07;]"
    m = isb.deserializeMatrix("[1.5,2,3;4,5,6.25;]")
    assert m.dtype == np.float32 and m.shape == (2, 3) and m[1, 2] == 6.25
    assert isb.deserializeMatrix("[7;8;9;]").shape == (3, 1)
    p = isb.parseMatrixStr("[1,0,0,0,0,1,0,0,0,0,1,0,0.5,0.25,2,1]")
    assert p.shape == (4, 4) and p[3, 1] == 0.25
    rig = synth.make_rig("cfg4")
    cams = isb.cameras_from_KR(rig.Ks, rig.Rs)
    cams[1].t[:] = [0.5, -1.25, 3.0]
    path = str(tmp_path / "cams.data")
    isb.serializeCameraParams(cams, path)
    line = open(path).read().splitlines()[1]
    assert line.count("@") == 5 and line.endswith(";]") and "@[0.5;-1.25;3;]@[" in line
    back = isb.deserializeCameraParams(path)
    assert len(back) == rig.n
    for a, b in zip(cams, back):
        # cams.data is lossy (6 significant digits, no setprecision) - same as the reference
        assert abs(a.focal - b.focal) <= 5e-6 * abs(a.focal) and np.allclose(list(a.R), list(b.R), atol=5e-6)
    assert list(back[1].t) == [0.5, -1.25, 3.0]
    # a second save/load round trip is a fixed point
    isb.serializeCameraParams(back, path)
    again = isb.deserializeCameraParams(path)
    assert all(list(x.R) == list(y.R) and x.focal == y.focal for x, y in zip(back, again))
    ipath = str(tmp_path / "indices.data")
    isb.serializeIndices([3, 0, 7], ipath)
    assert open(ipath).read() == "3\n0\n7\n"
    open(ipath, "a").write("\n9\n")
    assert isb.deserializeIndices(ipath) == [3, 0, 7, 9]
    with pytest.raises(isb.IsbError):
        isb.deserializeCameraParams(str(tmp_path / "missing.data"))

"""imwrite("result.jpg", result) (image_stitching.cpp:1228).

CPU: the numpy restatement of libjpeg's baseline encoder (oracle/jpeg_oracle.py) against cv2.imencode - the dependency the
reference calls - byte for byte.  GPU: isb_jpeg_encode through the C ABI against cv2.imencode (and, without cv2, against the
restatement), byte for byte: every MCU edge case (sizes that are not multiples of 8 / 16: replicated edges, dummy blocks),
flat / noisy / high-contrast content (long zero runs, ZRL, many 0xFF bytes to stuff), 16S input with values outside [0, 255],
a quality other than the default, and a composited panorama."""
import numpy as np
import pytest

from conftest import make_case, seam_masks_oracle
from oracle import jpeg_oracle as jo

try:
    import cv2
except Exception:  # noqa: BLE001
    cv2 = None


def images():
    rng = np.random.default_rng(11)
    out = []
    for k, (h, w) in enumerate([(16, 16), (1, 1), (8, 24), (24, 8), (40, 56), (33, 47), (17, 9), (100, 130), (121, 75), (7, 250), (64, 64)]):
        kind = k % 4
        if kind == 0:
            img = rng.integers(0, 256, (h, w, 3))
        elif kind == 1:
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([(xx * 3 + yy) % 256, (yy * 5) % 256, (xx + yy * 2) % 256], -1) + rng.integers(-4, 5, (h, w, 3))
        elif kind == 2:
            img = rng.integers(0, 2, (h, w, 3)) * 255
        else:
            img = np.full((h, w, 3), 200)
        out.append(np.clip(img, 0, 255).astype(np.uint8))
    return out


def reference_bytes(img, quality=95):
    if cv2 is not None:
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])
        assert ok
        return buf.tobytes()
    return jo.encode(img, quality)


@pytest.mark.skipif(cv2 is None, reason="cv2 pins the restatement")
def test_oracle_equals_cv2_imencode():
    for img in images():
        assert jo.encode(img) == cv2.imencode(".jpg", img)[1].tobytes(), img.shape
    img = images()[7]
    for q in (10, 50, 75, 100):
        assert jo.encode(img, q) == cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])[1].tobytes(), q


def test_oracle_stream_structure():
    """Without cv2: the marker sequence of the restatement and the quality-95 tables of jpeg_set_quality."""
    b = jo.encode(images()[4])
    assert b[:4] == bytes([0xFF, 0xD8, 0xFF, 0xE0]) and b[-2:] == bytes([0xFF, 0xD9]) and b[6:11] == b"JFIF\0"
    ql, qc = jo.quant_tables(95)
    assert list(ql[:8]) == [2, 1, 1, 2, 2, 4, 5, 6] and int(qc[63]) == 10
    markers, i = [], 2
    while b[i + 1] != 0xDA:
        markers.append(b[i + 1])
        i += 2 + ((b[i + 2] << 8) | b[i + 3])
    assert markers == [0xE0, 0xDB, 0xDB, 0xC0, 0xC4, 0xC4, 0xC4, 0xC4]


@pytest.mark.gpu
def test_gpu_jpeg_equals_reference_bytes():
    import image_stitching_b200 as isb
    for img in images():
        assert isb.imencode_jpg(img) == reference_bytes(img), img.shape
    img = images()[7]
    for q in (10, 75, 100):
        assert isb.imencode_jpg(img, q) == reference_bytes(img, q), q


@pytest.mark.gpu
def test_gpu_jpeg_16s_input_and_device_pointer():
    """blend() returns 16SC3; imwrite saturates it to 8U first (values below 0 and above 255 included)."""
    torch = pytest.importorskip("torch")
    import image_stitching_b200 as isb
    rng = np.random.default_rng(5)
    img16 = rng.integers(-40, 300, (75, 123, 3)).astype(np.int16)
    ref = reference_bytes(np.clip(img16, 0, 255).astype(np.uint8))
    assert isb.imencode_jpg(img16) == ref
    assert isb.imencode_jpg(torch.from_numpy(img16).cuda()) == ref


@pytest.mark.gpu
def test_gpu_jpeg_of_a_composited_panorama():
    """The output side end to end: compose on the device, encode the panorama where it lies, compare with imencode of the
    oracle's panorama."""
    torch = pytest.importorskip("torch")
    import image_stitching_b200 as isb
    from oracle import oracle as orc
    rig, imgs, gains, nb = make_case("cfg2", 8, 5)
    seams = seam_masks_oracle(rig)
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    c = isb.Composer(rig.warp, rig.scale, nb)
    c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
    x, y, w, h = c.dst_roi
    o8 = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    om = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    c.run([torch.from_numpy(im).cuda() for im in imgs], gains, seams, out=o8, out_mask=om)
    torch.cuda.synchronize()
    assert isb.imencode_jpg(o8) == reference_bytes(ref["result8"])

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import image_stitching_b200 as isb
        return isb.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: the product has no CPU fallback.
    pass


@pytest.fixture(scope="session")
def cv2_parity():
    cv2 = pytest.importorskip("cv2")
    from oracle import cv_reference as cvr
    cvr.set_parity_mode(True)
    return cv2


def make_case(name, div, nb=None, max_images=None, kind="texture", with_gains=True, with_seams=True):
    """Deterministic small rig + inputs shared by oracle, golden and GPU tests."""
    from image_stitching_b200 import synth
    rig = synth.make_rig(name, scale_div=div, max_images=max_images)
    imgs = [synth.make_image(i, rig.W, rig.H, kind) for i in range(rig.n)]
    gains = synth.make_gains(rig.n) if with_gains else None
    return rig, imgs, gains, (rig.nb if nb is None else nb)


def seam_masks_oracle(rig):
    """Seam masks through the C oracle's nearest warp (same construction as cv_reference.seam_masks_cv)."""
    from image_stitching_b200 import synth
    from oracle import oracle as orc
    src = synth.seam_source_mask(rig.W, rig.H)
    out = []
    for K, R in zip(rig.Ks, rig.Rs):
        Ks, ss = synth.seam_camera(K, rig.scale)
        _, m = orc.warp(rig.warp, ss, src, Ks, R, orc.NEAREST, 0)
        out.append(m)
    return out


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)

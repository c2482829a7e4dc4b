"""Deterministic synthetic rigs for the compositing hot path.

Everything here is integer / float64 numpy so that the same inputs are produced on
every machine: the golden vectors under tests/golden/ were generated from these
functions and are re-checked on the GPU box.

The rigs mirror BASELINE.json `configs`:
  cfg2  8 x 4000x3000, spherical, 5 bands          (the bench workload)
  cfg3  36 x 6000x4000, spherical, 7 bands, 360 deg
  cfg4  4 x 3840x2160, cylindrical, 5 bands (video rate, fixed cameras)
  cfg5  200 x 5472x3648, spherical, 8 bands (gigapixel mosaic)
`scale_div` shrinks the linear resolution (same angles) for parity-sized cases.

Camera convention follows the reference: K = [[f,0,W/2],[0,f,H/2],[0,0,1]]
(cv::detail::CameraParams::K(), image_stitching.cpp:1150-1151), R from
eulerAnglesToRotationMatrix(..., YXZ) (euler.h:135-300) evaluated in double and
cast to float32, warper scale = float32(f) (image_stitching.cpp:884-895,1116-1117).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

RIGS = {
    "cfg2": dict(W=4000, H=3000, pitches=[0.0], n_yaw=8, yaw0=-180.0, span=360.0, hfov=62.0,
                 warp="spherical", nb=5),
    "cfg3": dict(W=6000, H=4000, pitches=[-38.0, 0.0, 38.0], n_yaw=12, yaw0=-180.0, span=360.0, hfov=44.0,
                 warp="spherical", nb=7),
    "cfg4": dict(W=3840, H=2160, pitches=[0.0], n_yaw=4, yaw0=-100.0, span=200.0, hfov=70.0,
                 warp="cylindrical", nb=5),
    "cfg5": dict(W=5472, H=3648, pitches=[-18.0 + 4.0 * k for k in range(10)], n_yaw=20, yaw0=-50.0, span=100.0,
                 hfov=7.0, warp="spherical", nb=8),
}

JITTER = 0.01  # rad
GAIN_GRID = (5, 6)  # rows, cols of the BlocksGainCompensator gain map
SEAM_DIV = 8  # seam masks live at 1/8 of the compose resolution


def euler_yxz_to_R(x: float, y: float, z: float) -> np.ndarray:
    """three.js makeRotationFromEuler, order YXZ, as euler.h:174-193 + 289-297 (float64)."""
    a, b = math.cos(x), math.sin(x)
    c, d = math.cos(y), math.sin(y)
    e, f = math.cos(z), math.sin(z)
    ce, cf, de, df = c * e, c * f, d * e, d * f
    te = [0.0] * 16
    te[0] = ce + df * b
    te[4] = de * b - cf
    te[8] = a * d
    te[1] = a * f
    te[5] = a * e
    te[9] = -b
    te[2] = cf * b - de
    te[6] = df + ce * b
    te[10] = a * c
    return np.array([[te[0], te[4], te[8]], [te[1], te[5], te[9]], [te[2], te[6], te[10]]], dtype=np.float64)


@dataclass
class Rig:
    name: str
    warp: str
    nb: int
    W: int
    H: int
    scale: np.float32
    Ks: list = field(default_factory=list)  # float32 3x3
    Rs: list = field(default_factory=list)  # float32 3x3
    focal: float = 0.0

    @property
    def n(self) -> int:
        return len(self.Ks)


def make_rig(name: str, scale_div: int = 1, max_images: int | None = None) -> Rig:
    cfg = RIGS[name]
    W, H = cfg["W"] // scale_div, cfg["H"] // scale_div
    f = (W / 2.0) / math.tan(math.radians(cfg["hfov"]) / 2.0)
    rig = Rig(name=name, warp=cfg["warp"], nb=cfg["nb"], W=W, H=H, scale=np.float32(f), focal=f)
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]], dtype=np.float64).astype(np.float32)
    n_yaw = cfg["n_yaw"]
    for r, pitch in enumerate(cfg["pitches"]):
        for i in range(n_yaw):
            yaw = cfg["yaw0"] + cfg["span"] * (i + 0.5 * (r % 2)) / n_yaw
            ex = math.radians(pitch) + JITTER * math.cos(2 * i + r)
            ey = math.radians(yaw) + JITTER * math.sin(3 * i + r)
            ez = JITTER * math.sin(i + 2 * r)
            rig.Ks.append(K.copy())
            rig.Rs.append(euler_yxz_to_R(ex, ey, ez).astype(np.float32))
    if max_images is not None:
        rig.Ks, rig.Rs = rig.Ks[:max_images], rig.Rs[:max_images]
    return rig


def make_image(index: int, W: int, H: int, kind: str = "texture") -> np.ndarray:
    """8UC3 HWC image: integer-bilinear upsample (x32) of a seeded low-res field plus
    i.i.d. noise in [-10,10], clipped.  `checker`: hard 0/255 8-px checkerboard."""
    if kind == "checker":
        yy, xx = np.mgrid[0:H, 0:W]
        v = ((((xx >> 3) + (yy >> 3) + index) & 1) * 255).astype(np.uint8)
        return np.repeat(v[:, :, None], 3, axis=2).copy()
    rng = np.random.default_rng(1000 + index)
    hl, wl = H // 32 + 2, W // 32 + 2
    low = rng.integers(0, 256, (hl, wl, 3), dtype=np.int32)
    x = np.arange(W)
    xi, xa = x >> 5, (x & 31).astype(np.int32)[None, :, None]
    rows = low[:, xi, :] * (32 - xa) + low[:, xi + 1, :] * xa  # (hl, W, 3)
    out = np.empty((H, W, 3), dtype=np.uint8)
    CH = 512  # rows per chunk keeps the int32 temporaries small
    for y0 in range(0, H, CH):
        y = np.arange(y0, min(H, y0 + CH))
        yi, ya = y >> 5, (y & 31).astype(np.int32)[:, None, None]
        v = (rows[yi] * (32 - ya) + rows[yi + 1] * ya + 512) >> 10
        v += rng.integers(-10, 11, v.shape, dtype=np.int32)
        np.clip(v, 0, 255, out=v)
        out[y0:y0 + len(y)] = v
    return out


def make_gains(n: int) -> list:
    rng = np.random.default_rng(7)
    return [(1.0 + 0.05 * rng.standard_normal(GAIN_GRID)).astype(np.float32) for _ in range(n)]


def seam_source_mask(W: int, H: int) -> np.ndarray:
    """Seam-scale source mask: 255 in the central 75 % of the source columns."""
    ws, hs = W // SEAM_DIV, H // SEAM_DIV
    m = np.zeros((hs, ws), dtype=np.uint8)
    m[:, int(round(ws * 0.125)):int(round(ws * 0.875))] = 255
    return m


def seam_band_mask(w: int, h: int) -> np.ndarray:
    """w x h source-space mask: 255 in the central 75 % of the columns (a stand-in for the seam finder's output)."""
    m = np.zeros((h, w), dtype=np.uint8)
    m[:, int(round(w * 0.125)):int(round(w * 0.875))] = 255
    return m


def seam_camera(K: np.ndarray, scale: np.float32):
    """K and warper scale at seam resolution (image_stitching.cpp:973-983: focal, ppx, ppy
    scaled by seam_work_aspect; warper created with warped_image_scale * seam_work_aspect)."""
    Ks = K.astype(np.float32).copy()
    a = np.float32(1.0 / SEAM_DIV)
    Ks[0, 0] *= a
    Ks[0, 2] *= a
    Ks[1, 1] *= a
    Ks[1, 2] *= a
    return Ks, np.float32(np.float32(scale) * a)


def default_flow_setup(rig: Rig, compose_megapix: float = 0.4, seam_megapix: float = 0.1):
    """Scalars and cameras of the reference's DEFAULT invocation (image_stitching.cpp:53-55: work_megapix = -1, seam_megapix =
    0.1, compose_megapix = 0.4) for a rig of full-size frames: compose_scale (:1107-1108), the compose-scale cameras and warper
    scale (:1115-1127), the seam-scale cameras and warper scale (:973-983) and the compose size sz (:1130-1133)."""
    W, H = rig.W, rig.H
    compose_scale = min(1.0, math.sqrt(compose_megapix * 1e6 / (W * H)))
    seam_scale = min(1.0, math.sqrt(seam_megapix * 1e6 / (W * H)))
    work_aspect, seam_aspect = compose_scale / 1.0, seam_scale / 1.0  # work_scale = 1
    scale_c = np.float32(rig.scale) * np.float32(work_aspect)
    Kc, Ks = [], []
    for K in rig.Ks:
        Kd = np.eye(3)
        for (r, c) in ((0, 0), (1, 1), (0, 2), (1, 2)):
            Kd[r, c] = float(K[r, c]) * work_aspect
        Kc.append(Kd.astype(np.float32))
        Km = K.copy()
        swa = np.float32(seam_aspect)
        for (r, c) in ((0, 0), (0, 2), (1, 1), (1, 2)):
            Km[r, c] *= swa
        Ks.append(Km)
    rnd = lambda v: int(np.rint(v))  # noqa: E731  (cvRound)
    return dict(compose_scale=compose_scale, seam_scale=seam_scale, scale_c=scale_c, Kc=Kc, Ks=Ks,
                seam_warper_scale=np.float32(float(rig.scale) * seam_aspect), sz=(rnd(W * compose_scale), rnd(H * compose_scale)),
                seam_size=(rnd(W * seam_scale), rnd(H * seam_scale)))

"""Dev tool (GPU box): time of isb_jpeg_encode on a panorama-sized image against cv2.imencode (what imwrite calls), bytes compared.

  python tools/jpeg_time.py [--width 20912] [--height 2881] [--reps 5]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import image_stitching_b200 as isb  # noqa: E402
from image_stitching_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=20912)
    ap.add_argument("--height", type=int, default=2881)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    img = synth.make_image(1234, a.width, a.height)
    d = torch.from_numpy(img).cuda()
    w, h = a.width, a.height
    cap = w * h
    out_dev = torch.empty(cap, dtype=torch.uint8, device="cuda")
    out_pin = torch.empty(cap, dtype=torch.uint8).pin_memory()
    out_pag = np.empty(cap, np.uint8)
    out_pag[:] = 0
    n = C.c_size_t(0)
    L = isb.lib()

    def run(ptr):
        rc = L.isb_jpeg_encode(C.c_void_p(d.data_ptr()), w, h, C.c_size_t(w * 3), 0, 95, C.c_void_p(ptr), C.c_size_t(cap), C.byref(n))
        assert rc == 0, isb.last_error() if hasattr(isb, "last_error") else rc

    for name, ptr in (("device output", out_dev.data_ptr()), ("pinned host output", out_pin.data_ptr()), ("pageable host output", out_pag.ctypes.data)):
        run(ptr)
        ts = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(ptr)
            ts.append(time.perf_counter() - t0)
        print(f"{name}: {min(ts) * 1e3:.2f} ms best of {a.reps} ({w * h / 1e6 / min(ts):.0f} MP/s), {n.value} bytes", flush=True)
    try:
        import cv2
        t0 = time.perf_counter()
        ok, ref = cv2.imencode(".jpg", img)
        t = time.perf_counter() - t0
        print(f"cv2.imencode: {t * 1e3:.1f} ms ({w * h / 1e6 / t:.0f} MP/s), equal bytes: {ref.tobytes() == out_pag[:n.value].tobytes()}", flush=True)
    except ImportError:
        pass


if __name__ == "__main__":
    main()

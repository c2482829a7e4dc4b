"""Builds image_stitching_b200/libisb.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

  python -m image_stitching_b200.build [--force]

Host translation units that carry bit-exact float geometry are compiled with -ffp-contract=off;
device code with -fmad=false (the arithmetic contract of SURVEY.md Appendix A forbids FMA contraction).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libisb.so")
SOURCES = ["kernels.cu", "kernels_fast.cu", "simple_blend.cu", "crop.cu", "jpeg.cu", "engine.cu", "capi.cu", "geometry.cpp", "cam_io.cpp"]
HEADERS = ["kernels.cuh", "device_math.cuh", "glibc_math.cuh", "engine.hpp", "device_types.hpp", "geometry.hpp", "cam_io.hpp", "pose_math.hpp",
           os.path.join("..", "..", "include", "image_stitching.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libisb.so cannot be built (there is no CPU fallback)")


STAMP = SO + ".stamp"  # hash of the sources the binary was built from (file times do not survive a copy to the GPU box)


def source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(os.environ.get("ISB_NVCC_FLAGS", "").encode())
    return h.hexdigest()


def up_to_date() -> bool:
    if not (os.path.exists(SO) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return SO
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    common = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden,-Wall", "-I",
              os.path.join(HERE, "..", "include")]
    common += os.environ.get("ISB_NVCC_FLAGS", "").split()  # tuning experiments (-DISB_...), empty in production
    if verbose:
        common += ["-Xptxas", "-v"]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        cmd = [_nvcc()] + common + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src}\n{out}\n")
        elif verbose and out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libisb.so build failed")
    link = [_nvcc(), "-shared", "-o", SO] + objs + ["-cudart", "static", "-lpthread"]
    subprocess.check_call(link)
    with open(STAMP, "w") as fh:
        fh.write(source_hash() + "\n")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

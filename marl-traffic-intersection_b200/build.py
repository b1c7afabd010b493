"""In-tree build of the native libraries (no setuptools, no JIT cache: the .so files travel with the repo).

    python -m marl_traffic_intersection_b200.build        # or: python marl-traffic-intersection_b200/build.py

  csrc/libisx_b200.so        nvcc, sm_100a, -fmad=false   — the product
  csrc/libisx_math_host.so   g++                           — host build of the restated libm (CPU tests)
  csrc/libisx_host_units.so  g++                           — host build of the entity-level arithmetic (CPU tests)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unknown-pragmas,-Wno-maybe-uninitialized", "-shared",
]
GXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas", "-Wno-maybe-uninitialized",
             "-shared", "-x", "c++"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=CSRC)


def build(force: bool = False, verbose_ptxas: bool = False) -> None:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", "isx.h")]
    lib = os.path.join(CSRC, "libisx_b200.so")
    srcs = [os.path.join(CSRC, "isx_kernels.cu"), os.path.join(CSRC, "isx_api.cu"), os.path.join(CSRC, "isx_host_expand.cpp")]
    if force or _newer(lib, srcs + hdrs):
        _run([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["isx_kernels.cu", "isx_api.cu", "isx_host_expand.cpp", "-o", lib])
    for name in ("isx_math_host", "isx_host_units"):
        out = os.path.join(CSRC, f"lib{name}.so")
        src = os.path.join(CSRC, name + ".cpp")
        if force or _newer(out, [src] + hdrs):
            _run(["g++"] + GXX_FLAGS + [name + ".cpp", "-o", out, "-lpthread", "-lm"])


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)

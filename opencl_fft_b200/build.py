"""Build libb200fft.so (the CUDA engine + C ABI) and libcl_fft.so (the reference's C++ class interface on
top of it) in-tree with nvcc for sm_100a. No GPU is needed to build.

    python -m opencl_fft_b200.build            # build if sources are newer than the libraries
    python -m opencl_fft_b200.build --force
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
INCLUDE = os.path.join(ROOT, "include")
LIB_ENGINE = os.path.join(LIBDIR, "libb200fft.so")
LIB_CLASSES = os.path.join(LIBDIR, "libcl_fft.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-ccbin", "g++"] + os.environ.get("B2F_NVCC_FLAGS", "").split()


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*dirs: str) -> list[str]:
    out = []
    for d in dirs:
        if not os.path.isdir(d):
            continue
        for f in sorted(os.listdir(d)):
            p = os.path.join(d, f)
            if os.path.isfile(p) and f.endswith((".cu", ".cuh", ".h", ".cpp", ".inl")):
                out.append(p)
    return out


def _run(cmd: list[str]) -> None:
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + res.stdout + res.stderr)


def build(force: bool = False, verbose: bool = False) -> None:
    os.makedirs(LIBDIR, exist_ok=True)
    deps = _sources(CSRC, INCLUDE, os.path.join(INCLUDE, "CL"))
    if force or _newer(LIB_ENGINE, deps):
        cmd = [NVCC, *ARCH, *COMMON, "-shared", "-o", LIB_ENGINE, os.path.join(CSRC, "capi.cu")]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)
    if force or _newer(LIB_CLASSES, deps + [LIB_ENGINE]):
        cmd = [NVCC, *COMMON, "-shared", "-o", LIB_CLASSES, "-I", INCLUDE, os.path.join(CSRC, "cl_classes.cpp"),
               "-L", LIBDIR, "-lb200fft", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB_ENGINE, "and", LIB_CLASSES)

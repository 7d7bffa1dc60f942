"""In-tree build of the native code: libhf6d.so (CUDA, sm_100a) and the HoughForest CLI.

nvcc cross-compiles without a GPU, so this runs in the CPU-only container as well as on the GPU box.  Outputs live next
to the sources (git-ignored, but they travel with the gpurun snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhf6d.so")
CLI = os.path.join(HERE, "HoughForest")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off",
              "-DHF6D_MBAR_SPIN_LIMIT=16777216"]  # a barrier-protocol bug traps instead of hanging the GPU


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++, a relocated copy with broken spec files; prefer the system compiler
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources():
    out = [os.path.join(HERE, "..", "include", "hf6d.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or _newer(LIB, _sources()):
        cmd = [_nvcc(), "-ccbin", _host_cxx()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-shared", "-o", LIB, os.path.join(CSRC, "hf6d_api.cu")]
        subprocess.check_call(cmd)
    return LIB


def build_cli(force: bool = False) -> str:
    src = os.path.join(CSRC, "hough_forest_main.cpp")
    if not os.path.exists(src):
        return ""
    if force or _newer(CLI, [src, LIB]):
        build_lib()
        cmd = [_host_cxx(), "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(HERE, "..", "include"), "-o", CLI, src,
               "-L", HERE, "-lhf6d", "-lz", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return CLI


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_cli(force)


if __name__ == "__main__":
    build_all(force=True)
    print(LIB)

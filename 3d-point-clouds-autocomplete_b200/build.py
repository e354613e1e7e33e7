"""Build libhp_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python 3d-point-clouds-autocomplete_b200/build.py [--force] [--verbose]

Outputs: <pkg>/lib/libhp_b200.so (git-ignored; shipped to the GPU box by gpurun) and
per-source objects under <pkg>/build/.  No torch involved: the library only needs the CUDA
runtime, which is linked statically.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhp_b200.so")

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "-I" + os.path.join(REPO_ROOT, "include"), "-I" + CSRC,
              "-DHP_BUILDING_LIBRARY", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libhp_b200.so cannot be built (no CPU fallback exists)")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps_mtime() -> float:
    deps = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(REPO_ROOT, "include", "hp_b200.h")]
    return max(os.path.getmtime(p) for p in deps)


def _compile_one(src: str, force: bool, verbose: bool, extra) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(src), _deps_mtime())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    cmd = [_nvcc()] + ARCH_FLAGS + NVCC_FLAGS + list(extra) + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, force, verbose, extra_flags), srcs))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [_nvcc()] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                      extra_flags=["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else ())
    print(p)

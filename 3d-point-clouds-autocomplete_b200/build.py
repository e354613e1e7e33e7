"""Build libhp_b200.so (the C-ABI product library) and libhp_b200_bench.so (the same sources compiled once more
with -DHP_BENCH_BUILD: adds the measurement helpers of include/hp_b200_bench.h and the A/B environment switches; used by
bench.py and tools/ only) in-tree with nvcc for sm_100a.

    python 3d-point-clouds-autocomplete_b200/build.py [--force] [--verbose]

Outputs: <pkg>/lib/libhp_b200.so, <pkg>/lib/libhp_b200_bench.so (git-ignored; shipped to the GPU box by gpurun) and
per-source objects under <pkg>/build/ and <pkg>/build/bench/.  No torch involved: the libraries only need the CUDA
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
BENCH_LIB_PATH = os.path.join(LIB_DIR, "libhp_b200_bench.so")

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
    deps = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(REPO_ROOT, "include", "*.h"))
    return max(os.path.getmtime(p) for p in deps)


def _compile_one(src: str, force: bool, verbose: bool, extra, obj_dir: str = OBJ_DIR) -> str:
    obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(src), _deps_mtime())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    cmd = [_nvcc()] + ARCH_FLAGS + NVCC_FLAGS + list(extra) + ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj


def _link(lib_path: str, objs, force: bool, verbose: bool) -> None:
    if force or not os.path.exists(lib_path) or any(os.path.getmtime(o) > os.path.getmtime(lib_path) for o in objs):
        cmd = [_nvcc()] + ARCH_FLAGS + ["-shared", "-o", lib_path] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), bench: bool = True) -> str:
    """Compile every csrc/*.cu for sm_100a and link libhp_b200.so; with bench=True also libhp_b200_bench.so."""
    bench_dir = os.path.join(OBJ_DIR, "bench")
    os.makedirs(bench_dir, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = _sources()
    jobs = [(s, tuple(extra_flags), OBJ_DIR) for s in srcs]
    if bench:
        jobs += [(s, tuple(extra_flags) + ("-DHP_BENCH_BUILD",), bench_dir) for s in srcs]
    with cf.ThreadPoolExecutor(max_workers=min(16, len(jobs))) as ex:
        objs = list(ex.map(lambda j: _compile_one(j[0], force, verbose, j[1], j[2]), jobs))
    _link(LIB_PATH, objs[:len(srcs)], force, verbose)
    if bench:
        _link(BENCH_LIB_PATH, objs[len(srcs):], force, verbose)
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv,
                      extra_flags=(["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else []) + os.environ.get("HP_EXTRA_NVCC_FLAGS", "").split())
    print(p)

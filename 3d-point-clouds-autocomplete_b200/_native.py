"""ctypes binding of libhp_b200.so (include/hp_b200.h).

Host-side plumbing only: torch supplies device memory and the current stream, every
computation happens in the sm_100a kernels behind the C ABI.  There is NO fallback: if the
shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# HP_B200_LIB: load another build of the same sources instead (A/B measurements of compile-time variants, tools/ only)
LIB_PATH = os.environ.get("HP_B200_LIB") or os.path.join(_PKG_DIR, "lib", "libhp_b200.so")

HP_OK = 0
HP_ERR_INVALID_ARGUMENT, HP_ERR_CUDA, HP_ERR_UNSUPPORTED, HP_ERR_WORKSPACE = 1, 2, 3, 4

_lib = None
_lock = threading.Lock()

_vp = ctypes.c_void_p
_int = ctypes.c_int
_sz = ctypes.c_size_t
_ll = ctypes.c_longlong

# name -> (restype, argtypes).  Must list every function include/hp_b200.h declares
# (tests/test_abi_symbols.py cross-checks this table against the header).
SIGNATURES = {
    "hp_version": (_int, []),
    "hp_error_string": (ctypes.c_char_p, [_int]),
    "hp_last_error_message": (ctypes.c_char_p, []),
    "hp_nndistance": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_nndistance_ws": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hp_nndistancegrad": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_batch_pairwise_dist": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp]),
    "hp_chamfer_workspace_bytes": (_sz, [_int, _int, _int]),
    "hp_chamfer_forward": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hp_chamfer_backward": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_chamfer_inverse_ints": (_sz, [_int, _int, _int, _int]),
    "hp_chamfer_forward_inv": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hp_chamfer_step_supported": (_int, [_int, _int, _int]),
    "hp_chamfer_step": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hp_chamfer_backward_inv": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_nndistancegrad_inv": (_int, [_int, _int, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_approxmatch": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "hp_approxmatch_workspace_bytes": (_sz, [_int, _int, _int]),
    "hp_approxmatch_ws": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hp_matchcost": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "hp_matchcostgrad": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hp_emd_cost_workspace_bytes": (_sz, [_int, _int, _int]),
    "hp_emd_cost_pairs": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hp_emd_cost_pairs_fast": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "hp_target_network_num_weights": (_ll, [_int, ctypes.POINTER(_int), _int]),
    "hp_target_network_set_mode": (_int, [_int]),
    "hp_target_network_forward": (_int, [_int, _int, _int, ctypes.POINTER(_int), _int, _vp, _vp, _ll, _vp, _int, _vp]),
    "hp_target_network_backward_workspace_bytes": (_sz, [_int, _int, _int, ctypes.POINTER(_int), _int]),
    "hp_target_network_backward": (_int, [_int, _int, _int, ctypes.POINTER(_int), _int, _vp, _vp, _ll, _vp, _int, _vp, _vp,
                                          _vp, _sz, _vp]),
    "hp_pairwise_cd": (_int, [_int, _int, _int, _int, _vp, _vp, _int, _int, _vp, _vp]),
    "hp_pairwise_cd_pairs": (_int, [_ll, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
}

# include/hp_b200_bench.h: measurement helpers of libhp_b200_bench.so (bench.py / tools only; never on the product path)
BENCH_LIB_PATH = os.path.join(_PKG_DIR, "lib", "libhp_b200_bench.so")
BENCH_SIGNATURES = {
    "hp_measure_chamfer_ring_only": (_int, [_int, _int, _vp, _int, _vp, _vp, _sz, _vp]),
    "hp_measure_peak": (_int, [_int, _int, ctypes.POINTER(ctypes.c_double), _vp]),
    "hp_measure_set_trace": (_int, [_vp]),
}
_bench_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libhp_b200.so; raise loudly if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeLibraryMissing(
                    f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    f"or `python {os.path.join(_PKG_DIR, 'build.py')}`. There is no CPU or PyTorch fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the library is stale: loud by design
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != HP_OK:
        lib = load()
        msg = lib.hp_last_error_message().decode(errors="replace")
        kind = lib.hp_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {kind} (code {rc}): {msg}")


def load_bench() -> ctypes.CDLL:
    """Load libhp_b200_bench.so (the product sources + measurement helpers, built with -DHP_BENCH_BUILD)."""
    global _bench_lib
    if _bench_lib is not None:
        return _bench_lib
    with _lock:
        if _bench_lib is None:
            if not os.path.exists(BENCH_LIB_PATH):
                raise NativeLibraryMissing(f"{BENCH_LIB_PATH} is missing (build it with __graft_entry__.build())")
            lib = ctypes.CDLL(BENCH_LIB_PATH)
            for name, (res, args) in {**SIGNATURES, **BENCH_SIGNATURES}.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _bench_lib = lib
    return _bench_lib


def check_bench(rc: int, what: str) -> None:
    if rc != HP_OK:
        lib = load_bench()
        raise RuntimeError(f"{what} failed: {lib.hp_error_string(rc).decode()} (code {rc}): "
                           f"{lib.hp_last_error_message().decode(errors='replace')}")


def measure_peak(kind: int, iters: int = 4096, stream: int = 0) -> float:
    out = ctypes.c_double(0.0)
    check_bench(load_bench().hp_measure_peak(kind, iters, ctypes.byref(out), stream), "hp_measure_peak")
    return out.value

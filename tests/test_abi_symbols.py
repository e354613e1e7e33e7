"""CPU: the C-ABI library loads and exports every symbol include/hp_b200.h declares, and the
ctypes table covers exactly that set (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header="hp_b200.h"):
    text = open(os.path.join(REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"HP_API[^;(]*?\b(hp_\w+)\s*\(", text)))


def test_header_declares_the_reference_launchers():
    names = _declared()
    for must in ("hp_nndistance", "hp_nndistancegrad"):
        assert must in names


def test_library_exports_every_declared_symbol(hp):
    lib = ctypes.CDLL(hp._native.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/hp_b200.h but not exported"
    assert lib.hp_version() == 100


def test_ctypes_table_matches_header(hp):
    assert sorted(hp._native.SIGNATURES) == _declared()


def test_product_library_holds_no_measurement_code(hp):
    """hp_measure_* and the A/B environment switches live in libhp_b200_bench.so only (include/hp_b200_bench.h)."""
    bench_only = _declared("hp_b200_bench.h")
    assert bench_only == ["hp_measure_chamfer_ring_only", "hp_measure_peak", "hp_measure_set_trace"]
    assert sorted(hp._native.BENCH_SIGNATURES) == bench_only
    syms = subprocess.check_output(["nm", "-D", "--defined-only", hp._native.LIB_PATH], text=True)
    assert "hp_measure" not in syms
    blob = open(hp._native.LIB_PATH, "rb").read()
    for knob in (b"HP_RING_VARIANT", b"HP_NO_PDL", b"HP_NN_RING", b"HP_NN_VARIANT"):
        assert knob not in blob, knob
    bench = ctypes.CDLL(hp._native.BENCH_LIB_PATH)
    for name in bench_only + _declared():
        assert hasattr(bench, name), name


def test_no_torch_or_python_in_the_abi(hp):
    """The library must be loadable by any C host: it may not depend on libtorch / libpython."""
    out = subprocess.check_output(["ldd", hp._native.LIB_PATH], text=True)
    assert "torch" not in out and "python" not in out and "c10" not in out, out


def test_error_strings_without_gpu(hp):
    lib = hp._native.load()
    assert lib.hp_error_string(0) == b"ok"
    assert b"invalid" in lib.hp_error_string(1)
    # argument validation happens before any CUDA call
    rc = lib.hp_nndistance(-1, 1, None, 1, None, None, None, None, None, None)
    assert rc == 1 and b"negative" in lib.hp_last_error_message()
    rc = lib.hp_nndistance(2, 0, None, 5, None, None, None, None, None, None)
    assert rc == 1 and b"empty" in lib.hp_last_error_message()
    assert lib.hp_nndistance(0, 5, None, 5, None, None, None, None, None, None) == 0  # empty batch: no-op
    assert lib.hp_chamfer_workspace_bytes(32, 2048, 2048) >= 8 + 4 * 32 * 16

#!/usr/bin/env bash
python - <<'PY'
import importlib, sys
sys.path.insert(0, '.')
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
n = hp._native
print("FFMA  %.1f TFLOP/s" % (n.measure_peak(0, 8192) / 1e12))
print("mma.sync tf32 m16n8k8  %.1f TFLOP/s" % (n.measure_peak(6, 4096) / 1e12))
PY

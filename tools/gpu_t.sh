#!/usr/bin/env bash
set -uo pipefail
echo "== pytest gpu all"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== tn timing"; timeout 300 python tools/time_target_network.py 2>&1 | tail -6 | head -4
python - <<'PY'
import importlib, sys, torch
sys.path.insert(0, '.')
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
for B in (32, 64):
    st = hp.ChamferStepGraph(B, 2048, 2048, "cuda:0")
    for _ in range(5): st.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): st.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"chamfer step graph B={B}: {e0.elapsed_time(e1)/50*1e3:.1f} us (warm L2)")
PY

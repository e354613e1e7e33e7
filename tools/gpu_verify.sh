#!/usr/bin/env bash
# what the driver runs at round end on one GPU: smoke, the GPU tests, both bench arms
set -uo pipefail
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 200 gpurun_out/bench_reference.json
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench.json") if l.startswith("{")][-1])
r = d["roofline"]
print("value %.4e ms/step %.5f e2e %.4e launches %d | ring %.5f ms frac %.4f fwd %.4f fwd+bwd %.4f | eager %.4f ms" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], r["kernel_ms"], r["frac"], r["forward_frac"], r["fwd+bwd_frac"], d["eager_api"]["ms_per_step"]))
print({k: (round(v["ms"], 4) if isinstance(v, dict) and "ms" in v else v) for k, v in d["other_paths"].items()})
print(d["metrics_eval"])
PY

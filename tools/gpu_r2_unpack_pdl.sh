#!/usr/bin/env bash
# unpack kernel launched programmatically dependent on the ring kernel: tests of every caller, the probe, a short bench line
set -uo pipefail
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_chamfer_gpu.py tests/test_evaluation_gpu.py tests/test_metrics_gpu.py tests/test_native_binding.py -x -q -m gpu 2>&1 | tail -4
echo "== probe"; timeout 300 python tools/nn_flake_probe.py 3000 2>&1 | tail -2
echo "== bench (short)"; timeout 600 python bench.py --no-cpu-baseline --no-other-paths --no-metrics-eval > gpurun_out/r2_bench_unpack_pdl.json 2> gpurun_out/r2_bench_unpack_pdl.err
tail -2 gpurun_out/r2_bench_unpack_pdl.err
python - <<'P'
import json
d = json.load(open("gpurun_out/r2_bench_unpack_pdl.json"))
r = d["roofline"]
print("step ms", d["ms_per_step"], "ring", r["kernel_ms"], "forward (ring+unpack)", r["forward_ms"], "frac", r["forward_frac"], "bwd gather", r["bwd_kernel_ms"], "eager", d["eager_api"]["ms_per_step"])
print("e2e", d["e2e"]["ms_per_step"], "split2", d["e2e"]["split_in_two_halves_ms_per_step"])
P

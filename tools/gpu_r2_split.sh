#!/usr/bin/env bash
# the half-batch split of the host-IO Chamfer graph: its tests, then a short bench line (e2e split vs unsplit)
set -uo pipefail
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_target_network_gpu.py -x -q -m gpu -k "graph or out_buffers or pipeline" 2>&1 | tail -15
echo "== bench (short)"; timeout 600 python bench.py --no-cpu-baseline --no-other-paths --no-metrics-eval > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.err
tail -3 gpurun_out/r2_bench_split.err
python - <<'P'
import json
d = json.load(open("gpurun_out/r2_bench_split.json"))
print("value ms/step", d["ms_per_step"], "e2e", {k: d["e2e"][k] for k in ("ms_per_step", "host_io_split", "unsplit_ms_per_step", "with_gradients_d2h_ms_per_step", "pipelined_independent_steps_ms_per_step")})
P

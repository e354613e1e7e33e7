#!/usr/bin/env bash
# round 2: Chamfer step with the sectioned, ticket-driven tail: parity tests + quick bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu_info.csv
timeout 900 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2_chamfer_tests.log
cat gpurun_out/r2_chamfer_tests.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-other-paths --no-metrics-eval --no-cpu-baseline > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_quick.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("step ms", d["ms_per_step"], "ring ms", r["kernel_ms"], "fwd ms", r["forward_ms"], "bwd ms", r["bwd_kernel_ms"],
          "frac ring", r["frac"], "fwd+bwd frac", r["fwd+bwd_frac"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["unpipelined_ms_per_step"], "eager", d["eager_api"]["ms_per_step"])
except Exception as e:
    print("bench failed", e)
    print(open("gpurun_out/r2_bench_quick.err").read()[-3000:])
PY

#!/usr/bin/env bash
mkdir -p gpurun_out
HP_TAIL_TICKETS=1 timeout 120 python tools/chamfer_timeline.py > gpurun_out/r2_timeline3.txt 2>&1; head -14 gpurun_out/r2_timeline3.txt; tail -3 gpurun_out/r2_timeline3.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_gputests.log; cat gpurun_out/r2_gputests.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-other-paths --no-metrics-eval --no-cpu-baseline > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_quick.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("step ms", d["ms_per_step"], "ring ms", r["kernel_ms"], "fwd ms", r["forward_ms"], "bwd ms", r["bwd_kernel_ms"],
          "frac ring", r["frac"], "fwd+bwd frac", r["fwd+bwd_frac"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["with_gradients_d2h_ms_per_step"], d["e2e"]["pipelined_independent_steps_ms_per_step"], "eager", d["eager_api"]["ms_per_step"])
except Exception as e:
    print("bench failed", e)
    print(open("gpurun_out/r2_bench_quick.err").read()[-3000:])
PY

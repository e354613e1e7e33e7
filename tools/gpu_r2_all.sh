#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_gputests.log; cat gpurun_out/r2_gputests.log
( time timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err ) 2>&1 | tail -3
tail -5 gpurun_out/r2_bench.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err ) 2>&1 | tail -3
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench.json", "gpurun_out/r2_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "steps") if k in d}, "e2e", d.get("e2e", {}).get("value"))
        if "roofline" in d:
            r = d["roofline"]
            print(" roofline", {k: r[k] for k in ("frac", "kernel_ms", "forward_ms", "fwd+bwd_frac")})
            print(" metrics_eval", r.get("metrics_eval"))
            print(" cpu", d.get("cpu_baseline"), d.get("cpu_baseline_c1"))
            print(" other", json.dumps(d.get("other_paths"), indent=0)[:3000])
            print(" comparators", d.get("secondary_comparators"))
            print(" c4", d.get("c4_full_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest chamfer+tn"; timeout 900 python -m pytest tests/test_chamfer_gpu.py tests/test_target_network_gpu.py -x -q 2>&1 | grep -vE "^E   +\+" | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 60 --csv --log-file gpurun_out/launches_ring_warm.csv python tools/profile_chamfer.py 6 > /dev/null 2>&1
echo "warm (cache-control none):"; grep -E "nn_ring|nn_grad" gpurun_out/launches_ring_warm.csv | awk -F'","' '{print $5, $NF}' | tail -3
echo "== bench"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-other-paths --no-metrics-eval 2>gpurun_out/bench.err | tee gpurun_out/bench_v0.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e pairs/s  ms/step %.4f  e2e %.3e (%.1f us) eager %.1f us  fwd %.2fus bwd %.2fus frac %.3f fwd+bwd frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']*1e3, d['eager_api']['ms_per_step']*1e3, r['kernel_ms']*1e3, r['bwd_kernel_ms']*1e3, r['frac'], r['fwd+bwd_frac']))
"
tail -3 gpurun_out/bench.err

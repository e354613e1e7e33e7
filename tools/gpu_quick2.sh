#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest chamfer+tn"; timeout 900 python -m pytest tests/test_chamfer_gpu.py tests/test_target_network_gpu.py -x -q 2>&1 | tail -15
for ring in 1 0; do
echo "== bench HP_NN_RING=$ring"; HP_NN_RING=$ring timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_ring$ring.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e pairs/s  ms/step %.4f  e2e %.3e  fwd %.2fus bwd %.2fus frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms']*1e3, r['bwd_kernel_ms']*1e3, r['frac']))
"
tail -2 gpurun_out/bench.err
done


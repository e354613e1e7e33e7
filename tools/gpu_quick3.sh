#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out

echo "== pytest chamfer"; timeout 900 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | tail -3
echo "== bench"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_ring1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e pairs/s  ms/step %.4f  e2e %.3e  fwd %.2fus bwd %.2fus frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms']*1e3, r['bwd_kernel_ms']*1e3, r['frac']))
"
tail -2 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ring.csv python tools/profile_chamfer.py 6 > /dev/null 2>&1
grep -E "nn_ring|nn_grad" gpurun_out/launches_ring.csv | awk -F'","' '{print $5, $NF}' | tail -8
bash tools/gpu_profile.sh nn_ring_kernel prof_nn_ring2

#!/usr/bin/env bash
# gpurun --timeout 900 -- bash tools/gpu_quick.sh : chamfer tests + bench (no cpu baseline) + ncu full on the fwd kernel
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest chamfer"; timeout 600 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | tail -5
echo "== bench"; timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e pairs/s  ms/step %.4f  e2e %.3e  fwd %.2fus bwd %.2fus frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms']*1e3, r['bwd_kernel_ms']*1e3, r['frac']))
print({k:round(v,2) for k,v in r.items() if 'tflops' in k}, d['clocks'])
"
tail -2 gpurun_out/bench.err
if [ "${1:-}" != "noprof" ]; then bash tools/gpu_profile.sh "${1:-nn_fwd}" "${2:-prof_nn_fwd}"; fi

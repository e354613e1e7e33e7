#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest emd"; timeout 900 python -m pytest tests/test_emd_gpu.py -x -q 2>&1 | tail -15
echo "== pytest chamfer"; timeout 600 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | tail -3
for v in 0 5 6 7 8; do
  echo "== chamfer variant $v"
  HP_NN_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e pairs/s  ms/step %.4f  e2e %.3e  fwd %.2fus bwd %.2fus frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms']*1e3, r['bwd_kernel_ms']*1e3, r['frac']))
"
done
echo "== emd timing"; timeout 300 python tools/time_emd.py 2>&1 | tail -12

#!/usr/bin/env bash
# gpurun -- bash tools/gpu_sweep.sh : forward-variant sweep (HP_NN_VARIANT) + backward profile
set -uo pipefail
mkdir -p gpurun_out
for v in 0 1 2 3 4; do
  echo "== variant $v"
  HP_NN_VARIANT=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e pairs/s  ms/step %.4f  e2e %.3e  fwd %.2fus bwd %.2fus frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['kernel_ms']*1e3, r['bwd_kernel_ms']*1e3, r['frac']))
"
done
HP_NN_VARIANT=0 timeout 300 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | tail -2
bash tools/gpu_profile.sh nn_grad prof_nn_grad
bash tools/gpu_profile.sh nn_fwd prof_nn_fwd_v0

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | grep -vE "^E   +\+" | tail -6
for v in 0 2; do
HP_RING_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 40 --csv --log-file gpurun_out/l_$v.csv python tools/profile_chamfer.py 6 > /dev/null 2>&1
echo "variant $v:"; grep -E "nn_ring|nn_grad" gpurun_out/l_$v.csv | awk -F'","' '{print $5, $NF}' | tail -4
done

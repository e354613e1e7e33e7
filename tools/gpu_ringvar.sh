#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
for v in 0 1 2; do
HP_RING_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 40 --csv --log-file gpurun_out/l_$v.csv python tools/profile_chamfer.py 6 > /dev/null 2>&1
echo "variant $v:"; grep -E "nn_ring_kernel" gpurun_out/l_$v.csv | awk -F'","' '{print $5, $NF}' | tail -3
done

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 40 --csv --log-file gpurun_out/l_0.csv python tools/profile_chamfer.py 6 > /dev/null 2>&1
grep -E "nn_ring|nn_grad" gpurun_out/l_0.csv | awk -F'","' '{print $5, $NF}' | tail -3

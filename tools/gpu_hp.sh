#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | grep -vE "^E   +\+" | tail -3
timeout 300 python tools/time_hot_path.py 2>&1 | tail -7
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/l_hp.csv python tools/time_hot_path.py > /dev/null 2>&1
grep -E "nn_ring|nn_grad|tn_" gpurun_out/l_hp.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -12

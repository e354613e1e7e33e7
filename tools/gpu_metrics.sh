#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest metrics"; timeout 900 python -m pytest tests/test_metrics_gpu.py -x -q 2>&1 | tail -15
echo "== time"; timeout 300 python tools/time_metrics.py 1000 1000 2>&1 | tail -3
DRIVER=tools/time_metrics.py bash tools/gpu_profile.sh pairwise_cd prof_pairwise_cd 96 96

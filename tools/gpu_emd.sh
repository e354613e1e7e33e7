#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest emd+metrics"; timeout 900 python -m pytest tests/test_emd_gpu.py tests/test_metrics_gpu.py -x -q 2>&1 | tail -4
echo "== emd timing"; timeout 300 python tools/time_emd.py 2>&1 | tail -8

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_metrics_gpu.py tests/test_evaluation_gpu.py -x -q 2>&1 | tail -4
timeout 600 python tools/time_metrics.py 2>&1 | tail -8

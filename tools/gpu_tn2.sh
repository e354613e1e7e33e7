#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest target network"; timeout 900 python -m pytest tests/test_target_network_gpu.py -x -q 2>&1 | tail -3
echo "== timing"; timeout 300 python tools/time_target_network.py 2>&1 | tail -6

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== pytest target network"; timeout 900 python -m pytest tests/test_target_network_gpu.py -x -q 2>&1 | tail -15
echo "== timing"; timeout 300 python tools/time_target_network.py 2>&1 | tail -8
DRIVER=tools/time_target_network.py bash tools/gpu_profile.sh tn_forward prof_tn_fwd 2
DRIVER=tools/time_target_network.py bash tools/gpu_profile.sh tn_backward prof_tn_bwd 2

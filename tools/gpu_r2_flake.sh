#!/usr/bin/env bash
# the GPU suite as the FIRST CUDA process of a fresh box (the condition of DESIGN.md 8's open issue), full output kept, then the probe
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
echo "== pytest -m gpu (first CUDA process)"; timeout 900 python -m pytest tests/ -x -q -m gpu -rw 2>&1 | tail -40 | tee gpurun_out/r2_gputests_first.log | tail -15
echo "== probe"; timeout 300 python tools/nn_flake_probe.py 6000 > gpurun_out/nn_flake_probe.txt 2>&1; tail -5 gpurun_out/nn_flake_probe.txt

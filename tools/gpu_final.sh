#!/usr/bin/env bash
# final single-GPU record: tests, bench (both arms), launch list, full ncu captures, sanitizer
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_reference.json
bash tools/gpu_profiles_refresh.sh
SAN_TOOLS="${SAN_TOOLS:-memcheck racecheck}" bash tools/gpu_sanitize.sh

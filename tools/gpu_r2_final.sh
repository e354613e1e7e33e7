#!/usr/bin/env bash
# round-2 final single-GPU record: all GPU tests, both bench arms, the ncu launch list of the bench command
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu -rw 2>&1 | tail -12 | tee gpurun_out/r2_gputests.log | tail -3
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r2_bench_reference.json
echo "== bench"; ( time timeout 1200 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err ) 2>&1 | tail -3; tail -2 gpurun_out/r2_bench.err
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths --no-metrics-eval > gpurun_out/r02_bench_under_ncu.log 2>&1
grep -c . gpurun_out/r02_launches.csv

#!/usr/bin/env bash
# Run on the GPU box via: gpurun --timeout 1500 -- bash tools/gpu_check.sh
# smoke + GPU parity tests + golden generation + bench + ncu launch list.
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -5
echo "== golden (reference ext)"; python tests/golden/make_golden_gpu.py 2>&1 | tail -2
cp gpurun_out/gpu_reference_ext.npz tests/golden/ 2>/dev/null
echo "== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -25
echo "== bench"; timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 600 gpurun_out/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches.csv

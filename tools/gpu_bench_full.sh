#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
echo "== bench full (N=1)"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_reference.json

#!/usr/bin/env bash
# gpurun --timeout 2400 -- bash tools/gpu_round1.sh : everything the round needs in one box visit.
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 800 gpurun_out/bench_reference.json
echo "== emd timing"; timeout 300 python tools/time_emd.py 2>&1 | tail -8
echo "== metrics timing"; timeout 300 python tools/time_metrics.py 1000 1000 2>&1 | tail -3
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches.csv
bash tools/gpu_profile.sh nn_fwd prof_nn_fwd
bash tools/gpu_profile.sh nn_grad prof_nn_grad
DRIVER=tools/time_metrics.py bash tools/gpu_profile.sh pairwise_cd prof_pairwise_cd 96 96

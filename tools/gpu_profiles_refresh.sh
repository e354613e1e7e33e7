#!/usr/bin/env bash
# refresh the ncu evidence under gpurun_out/ (copied to profiles/ afterwards)
set -uo pipefail
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths --no-metrics-eval > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches.csv
bash tools/gpu_profile.sh nn_ring_kernel prof_nn_ring 8 step
bash tools/gpu_profile.sh nn_ring_finish prof_nn_finish 8 step
bash tools/gpu_profile.sh nn_ring_unpack prof_nn_unpack
bash tools/gpu_profile.sh nn_grad_gather prof_nn_gather
DRIVER=tools/time_emd.py bash tools/gpu_profile.sh emd_pass prof_emd

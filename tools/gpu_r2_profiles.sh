#!/usr/bin/env bash
# round-2 ncu / sanitizer evidence -> gpurun_out/ (summaries are copied to profiles/ afterwards)
set -uo pipefail
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-paths --no-metrics-eval > gpurun_out/r02_bench_under_ncu.log 2>&1
grep -c . gpurun_out/r02_launches.csv
bash tools/gpu_profile.sh nn_ring_kernel r02_prof_nn_ring 8 step
bash tools/gpu_profile.sh nn_ring_tail r02_prof_nn_tail 8 step
bash tools/gpu_profile.sh nn_ring_unpack r02_prof_nn_unpack
DRIVER=tools/time_emd.py bash tools/gpu_profile.sh emd_pass r02_prof_emd
DRIVER=tools/time_emd.py bash tools/gpu_profile.sh emd_fused31 r02_prof_emd_fused31
DRIVER=tools/sanitizer_driver.py bash tools/gpu_profile.sh batch_pairwise_dist r02_prof_bpd
bash tools/gpu_sanitize.sh
cp gpurun_out/sanitizer_memcheck.log gpurun_out/r02_sanitizer_memcheck.log; cp gpurun_out/sanitizer_racecheck.log gpurun_out/r02_sanitizer_racecheck.log

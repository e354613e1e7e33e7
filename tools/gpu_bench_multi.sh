#!/usr/bin/env bash
# gpurun --gpus N -- bash tools/gpu_bench_multi.sh N
set -uo pipefail
N="${1:-2}"
mkdir -p gpurun_out
nvidia-smi -L
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus "$N" --steps 100 --warmup 5 > "gpurun_out/bench_n$N.json" 2> "gpurun_out/bench_n$N.err"
tail -c 2500 "gpurun_out/bench_n$N.json"; tail -5 "gpurun_out/bench_n$N.err"
echo "== metrics equivalence G=1 vs G=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 \
    tools/check_metrics_multi.py 2>&1 | tail -8

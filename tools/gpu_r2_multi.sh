#!/usr/bin/env bash
# usage: gpurun --gpus N -- bash tools/gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_metrics_multirank_gpu.py -q -s 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n$N.json") if l.startswith("{")][-1])
print("n_gpus", d["n_gpus"], "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
print(d["roofline"]["metrics_eval"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus $N --steps 5 --warmup 1 | tail -1 | cut -c1-300

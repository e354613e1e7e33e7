#!/usr/bin/env bash
# gpurun --timeout 1200 -- bash tools/gpu_profile.sh <kernel-regex> <out-name> [driver.py args...]
set -uo pipefail
mkdir -p gpurun_out
REGEX="${1:-nn_fwd}"; OUT="${2:-prof}"; shift 2 || true
DRIVER="${DRIVER:-tools/profile_chamfer.py}"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${REGEX}" -s 2 -c 2 -f -o "gpurun_out/${OUT}" \
    python "$DRIVER" "$@" > "gpurun_out/${OUT}.log" 2>&1
tail -3 "gpurun_out/${OUT}.log"; ls -la gpurun_out/${OUT}.ncu-rep

#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
for tool in ${SAN_TOOLS:-memcheck racecheck}; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_driver.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -6 gpurun_out/sanitizer_$tool.log
done

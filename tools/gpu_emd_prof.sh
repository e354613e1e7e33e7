#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:emd_pass" -s 30 -c 3 -f -o gpurun_out/prof_emd python tools/time_emd.py > gpurun_out/prof_emd.log 2>&1
tail -2 gpurun_out/prof_emd.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 120 --csv --log-file gpurun_out/launches_emd.csv python tools/time_emd.py > /dev/null 2>&1
awk -F'","' 'NR>2{print $5, $9, $8, $NF}' gpurun_out/launches_emd.csv | sed 's/hp:://' | cut -c1-150 | sed -n 1,45p

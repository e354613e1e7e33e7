#!/usr/bin/env bash
mkdir -p gpurun_out
{
python tools/chamfer_timeline.py
HP_NO_PDL=1 python tools/chamfer_timeline.py | head -1
HP_RING_VARIANT=20 python tools/chamfer_timeline.py | tail -1
} > gpurun_out/r2_timeline.txt 2>&1
cat gpurun_out/r2_timeline.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_timeline.csv python tools/chamfer_timeline.py > /dev/null 2>&1
grep -o '"nn_ring[a-z_]*[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/r2_launches_timeline.csv | awk -F'","' '{print $1, $NF}' | sort | uniq -c | sort -rn | head -12

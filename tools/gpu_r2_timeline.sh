#!/usr/bin/env bash
# every python call under its own timeout: a kernel that waits for a ticket that never comes must not eat the GPU budget
mkdir -p gpurun_out
{
echo "===== tickets (default)"; HP_TAIL_TICKETS=1 timeout 120 python tools/chamfer_timeline.py || echo "FAILED/TIMEOUT rc=$?"
echo "===== whole-grid wait";   HP_TAIL_TICKETS=0 timeout 120 python tools/chamfer_timeline.py || echo "FAILED/TIMEOUT rc=$?"
echo "===== whole-grid wait, 5 CTAs/SM";   HP_TAIL_TICKETS=0 HP_RING_VARIANT=1 HP_TIMELINE_BRIEF=1 timeout 120 python tools/chamfer_timeline.py || echo "FAILED/TIMEOUT rc=$?"
echo "===== tickets, 5 CTAs/SM";   HP_TAIL_TICKETS=1 HP_RING_VARIANT=1 HP_TIMELINE_BRIEF=1 timeout 120 python tools/chamfer_timeline.py || echo "FAILED/TIMEOUT rc=$?"
} > gpurun_out/r2_timeline2.txt 2>&1
cat gpurun_out/r2_timeline2.txt

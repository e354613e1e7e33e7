#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_metrics_reference_parity_gpu.py tests/test_emd_gpu.py tests/test_metrics_gpu.py -q -s 2>&1 | grep "EMDREL\|FLIPS\|passed\|failed\|Error\|error" | head -20
timeout 300 python tools/time_emd.py 2>&1 | tail -12
HP_TAIL_TICKETS=1 timeout 120 python tools/chamfer_timeline.py | head -12

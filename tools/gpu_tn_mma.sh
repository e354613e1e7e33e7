#!/bin/bash
# TargetNetwork 3xTF32 tensor-core path: parity tests + timing of both modes (gpurun)
mkdir -p gpurun_out
python -m pytest tests/test_target_network_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/tn_mma_tests.log
python tools/time_target_network.py 20 > gpurun_out/tn_mma_time.log 2>&1
HP_TN_MODE=fp32 python tools/time_target_network.py 20 > gpurun_out/tn_fp32_time.log 2>&1
tail -20 gpurun_out/tn_mma_tests.log; cat gpurun_out/tn_mma_time.log; cat gpurun_out/tn_fp32_time.log

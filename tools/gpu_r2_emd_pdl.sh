#!/usr/bin/env bash
# EMD auction chain with programmatic dependent launches: parity tests, then A/B timing (bench library, HP_EMD_NO_PDL)
set -uo pipefail
mkdir -p gpurun_out
echo "== EMD tests"; timeout 900 python -m pytest tests/test_emd_gpu.py tests/test_metrics_gpu.py tests/test_metrics_reference_parity_gpu.py -x -q -m gpu 2>&1 | tail -4
BL=$PWD/3d-point-clouds-autocomplete_b200/lib/libhp_b200_bench.so
{
echo "== product library (PDL)"; timeout 300 python tools/time_emd.py 2>&1 | grep -v Warning
echo "== bench library, HP_EMD_NO_PDL=1"; HP_B200_LIB=$BL HP_EMD_NO_PDL=1 timeout 300 python tools/time_emd.py 2>&1 | grep -v Warning
echo "== bench library, HP_EMD_NO_PDL=0"; HP_B200_LIB=$BL HP_EMD_NO_PDL=0 timeout 300 python tools/time_emd.py 2>&1 | grep -v Warning
} | tee gpurun_out/r2_emd_pdl.txt

#!/bin/bash
# time the TargetNetwork kernels of every variant library under lib/variants (gpurun)
for v in 3d-point-clouds-autocomplete_b200/lib/variants/*.so; do
  echo "== $v"; HP_B200_LIB=$PWD/$v timeout 120 python tools/time_target_network.py 20 2>&1 | grep "^B=64"
done

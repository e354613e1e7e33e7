#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
bash tools/gpu_profile.sh nn_ring_unpack prof_nn_unpack

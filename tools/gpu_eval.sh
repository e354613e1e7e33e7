#!/usr/bin/env bash
set -uo pipefail
timeout 600 python -m pytest tests/test_evaluation_gpu.py -x -q 2>&1 | grep -vE "^E   +\+" | tail -15

#!/usr/bin/env bash
timeout 900 python -m pytest tests/test_chamfer_gpu.py -x -q 2>&1 | grep -vE "^E   +\+   " | tail -25

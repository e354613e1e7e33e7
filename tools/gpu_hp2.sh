#!/usr/bin/env bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | grep -vE "^E   +\+   " | tail -22

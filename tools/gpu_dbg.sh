#!/usr/bin/env bash
timeout 300 python tools/debug_emd.py 2>&1 | tail -5

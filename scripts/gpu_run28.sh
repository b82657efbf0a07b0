#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 bash scripts/sanitize.sh > $O/r28_sanitize.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" $O/r28_sanitize.log
timeout -s KILL 600 python scripts/soak.py 300 5000 > $O/r28_soak.log 2>&1; tail -3 $O/r28_soak.log

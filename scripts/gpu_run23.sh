#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -k "test_random_event_stream and rounds" -x -q > $O/r23_san.log 2>&1; grep -n "=========" $O/r23_san.log | head -40

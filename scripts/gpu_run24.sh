#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r22_ab.txt
timeout -s KILL 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -k "test_random_event_stream and rounds" -x -q > $O/r23_san.log 2>&1; tail -3 $O/r23_san.log
bash scripts/gpu_run22.sh

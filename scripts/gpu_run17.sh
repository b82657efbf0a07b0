#!/bin/bash
# range profile of one live step (concurrent kernels): how busy are DRAM, L2 and the SMs while it runs?
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 600 ncu --replay-mode range --set full --clock-control none -f -o $O/r17_range python scripts/range_step.py > $O/r17_range.out 2>&1; tail -5 $O/r17_range.out
ls -la $O/r17_range.ncu-rep

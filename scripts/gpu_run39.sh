#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_green scripts/ubench_green.cu -lcuda > gpurun_out/r39_green.txt 2>&1
timeout -s KILL 200 scripts/ubench_green >> gpurun_out/r39_green.txt 2>&1; cat gpurun_out/r39_green.txt

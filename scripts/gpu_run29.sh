#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_fnv scripts/ubench_fnv.cu > gpurun_out/r29_fnv.txt 2>&1
timeout -s KILL 120 scripts/ubench_fnv >> gpurun_out/r29_fnv.txt 2>&1; cat gpurun_out/r29_fnv.txt

#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_lcp scripts/ubench_lcp.cu > gpurun_out/r20_lcp.txt 2>&1
timeout -s KILL 120 scripts/ubench_lcp >> gpurun_out/r20_lcp.txt 2>&1; cat gpurun_out/r20_lcp.txt

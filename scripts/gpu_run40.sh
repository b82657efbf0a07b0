#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 600 python bench.py > $O/r40_bench.json 2> $O/r40_bench.err; tail -c 300 $O/r40_bench.json; tail -2 $O/r40_bench.err

#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/r26_pytest.log 2>&1; tail -3 $O/r26_pytest.log
timeout -s KILL 600 python bench.py > $O/r26_bench.json 2> $O/r26_bench.err; tail -c 400 $O/r26_bench.json

#!/bin/bash
# 1-GPU session: kernel timelines of live steps (three batch sizes), then the GPU suite and the default bench on the final code
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
for n in 1048576 65536 524288; do
  timeout -s KILL 300 python scripts/timeline.py 10000000 $n $O/r11_timeline_$n.json > $O/r11_timeline_$n.out 2>&1; tail -3 $O/r11_timeline_$n.out
done
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/r11_pytest.log 2>&1; tail -5 $O/r11_pytest.log
timeout -s KILL 600 python bench.py > $O/r11_bench.json 2> $O/r11_bench.err; tail -c 1500 $O/r11_bench.json

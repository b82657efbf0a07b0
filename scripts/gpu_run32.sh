#!/bin/bash
# member list (representatives + followers in live-list order): parity, then A/B against the previous build, with kernel durations
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r32_ab.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q > $O/r32_pytest.log 2>&1; tail -2 $O/r32_pytest.log
for n in 1048576 524288; do
echo "== member list, $n prompts" >> $O/r32_ab.txt
timeout -s KILL 300 python scripts/ab_step.py 10000000 $n base >> $O/r32_ab.txt 2>&1
echo "== previous build" >> $O/r32_ab.txt
KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/prev/libkvidx.so timeout -s KILL 300 python scripts/ab_step.py 10000000 $n base >> $O/r32_ab.txt 2>&1
done
cat $O/r32_ab.txt
timeout -s KILL 300 python scripts/timeline.py 10000000 1048576 $O/r32_tl.json > $O/r32_tl.out 2>&1; tail -1 $O/r32_tl.out

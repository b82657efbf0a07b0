#!/bin/bash
# A/B: ring depth of kernel G (chunks in flight per warp) and the blocks a partial follower walks alone per round
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r35_ab.txt
echo "== default (ring 4 = 2 chunks in flight, 3 detach blocks)" >> $O/r35_ab.txt
timeout -s KILL 300 python scripts/ab_step.py 10000000 1048576 base >> $O/r35_ab.txt 2>&1
for v in RING5 RING6 DETACH1 DETACH2 DETACH6; do
  echo "== $v" >> $O/r35_ab.txt
  KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/$v/libkvidx.so timeout -s KILL 300 python scripts/ab_step.py 10000000 1048576 base >> $O/r35_ab.txt 2>&1
done
cat $O/r35_ab.txt

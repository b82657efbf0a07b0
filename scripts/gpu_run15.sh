#!/bin/bash
# A/B: prompts per warp tile of kernel G (32 / 16 / 8) x stream arrangement
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
C="base KVIDX_GROUP_SERIAL=2 KVIDX_GROUP_SERIAL=1,KVIDX_ROUNDS_PARTS=2,KVIDX_GROUP_SERIAL_GRID=3 KVIDX_GROUP_SERIAL=1,KVIDX_ROUNDS_PARTS=4,KVIDX_GROUP_SERIAL_GRID=3 KVIDX_ROUNDS_PARTS=4 KVIDX_GROUP_SERIAL=2,KVIDX_ROUNDS_GRID=3,4,2,4,4"
for t in 16 8; do
  echo "== tile $t" >> $O/r15_ab.txt
  KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/t$t/libkvidx.so timeout -s KILL 600 python scripts/ab_step.py 10000000 1048576 $C >> $O/r15_ab.txt 2>&1
done
cat $O/r15_ab.txt
KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/t8/libkvidx.so KVIDX_GROUP_SERIAL=1 KVIDX_ROUNDS_PARTS=2 KVIDX_GROUP_SERIAL_GRID=3 timeout -s KILL 300 python scripts/timeline.py 10000000 1048576 $O/r15_tl_t8ser2.json > $O/r15_tl.out 2>&1; tail -2 $O/r15_tl.out

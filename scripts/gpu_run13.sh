#!/bin/bash
# A/B: CTA size of the short kernels x stream priorities x kernel G's grid
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python scripts/ab_step.py 10000000 1048576 base KVIDX_SMALL_CTA=128 KVIDX_SMALL_CTA=64 KVIDX_SMALL_CTA=32 KVIDX_GROUP_SERIAL=2 KVIDX_GROUP_SERIAL=2,KVIDX_SMALL_CTA=64 \
   KVIDX_GROUP_SERIAL=2,KVIDX_SMALL_CTA=64,KVIDX_ROUNDS_GRID=3,4,2,4,4 KVIDX_SMALL_CTA=64,KVIDX_ROUNDS_GRID=3,4,2,4,4 KVIDX_SMALL_CTA=64,KVIDX_ROUNDS_GRID=1,4,2,4,4 \
   KVIDX_GROUP_SERIAL=2,KVIDX_SMALL_CTA=64,KVIDX_ROUNDS_PARTS=12 KVIDX_GROUP_SERIAL=2,KVIDX_SMALL_CTA=64,KVIDX_ROUNDS_PARTS=6 > $O/r13_ab.txt 2>&1
cat $O/r13_ab.txt
timeout -s KILL 300 python scripts/timeline.py 10000000 1048576 $O/r13_tl_base.json > $O/r13_tl_base.out 2>&1; tail -2 $O/r13_tl_base.out
KVIDX_GROUP_SERIAL=2 KVIDX_SMALL_CTA=64 timeout -s KILL 300 python scripts/timeline.py 10000000 1048576 $O/r13_tl_prio64.json > $O/r13_tl_prio64.out 2>&1; tail -2 $O/r13_tl_prio64.out

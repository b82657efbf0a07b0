#!/bin/bash
# A/B: few large parts, kernel G of the parts one after the other (full bandwidth each), the other part's short kernels beside it
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
C=""
for p in 2 3 4; do for g in 2 3; do C="$C KVIDX_GROUP_SERIAL=1,KVIDX_ROUNDS_PARTS=$p,KVIDX_GROUP_SERIAL_GRID=$g"; done; done
timeout -s KILL 900 python scripts/ab_step.py 10000000 1048576 base KVIDX_GROUP_SERIAL=2 KVIDX_ROUNDS_PARTS=2 KVIDX_ROUNDS_PARTS=4 $C KVIDX_GROUP_SERIAL=2,KVIDX_ROUNDS_PARTS=4 KVIDX_GROUP_SERIAL=2,KVIDX_ROUNDS_PARTS=6 KVIDX_GROUP_SERIAL=2,KVIDX_ROUNDS_PARTS=12 > $O/r14_ab.txt 2>&1
cat $O/r14_ab.txt
KVIDX_GROUP_SERIAL=1 KVIDX_ROUNDS_PARTS=2 KVIDX_GROUP_SERIAL_GRID=3 timeout -s KILL 300 python scripts/timeline.py 10000000 1048576 $O/r14_tl_ser2.json > $O/r14_tl_ser2.out 2>&1; tail -2 $O/r14_tl_ser2.out

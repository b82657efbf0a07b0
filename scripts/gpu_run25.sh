#!/bin/bash
# where does the warp-per-prompt round configuration beat the fused kernel (small side) and the lane-per-prompt rounds (large side)?
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r25_ab.txt
for n in 4096 8192 16384 24576 32768 49152; do
  echo "== $n prompts" >> $O/r25_ab.txt
  timeout -s KILL 300 python scripts/ab_step.py 10000000 $n base KVIDX_ROUNDS_MIN=2048,KVIDX_ROUNDS_WARP_MAX=1000000 KVIDX_ROUNDS_MIN=2048,KVIDX_ROUNDS_WARP=0 KVIDX_COOP_MAX=100000 >> $O/r25_ab.txt 2>&1
done
cat $O/r25_ab.txt

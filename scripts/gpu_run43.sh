#!/bin/bash
# lane-per-prompt rounds: kernel H with 3 stages / 128-thread CTAs vs 2 stages / 256
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r43_ab.txt
for n in 65536 131072 262144; do
  echo "== $n prompts" >> $O/r43_ab.txt
  timeout -s KILL 300 python scripts/ab_step.py 10000000 $n base KVIDX_ROUNDS_LANE_STAGES=3 >> $O/r43_ab.txt 2>&1
done
cat $O/r43_ab.txt

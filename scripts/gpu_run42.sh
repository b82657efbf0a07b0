#!/bin/bash
# where does the class pipeline take over from the per-prompt rounds now?  (KVIDX_CLASSES_MIN, 393216 so far)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r42_ab.txt
for n in 131072 196608 262144 327680 393216; do
  echo "== $n prompts: class pipeline forced, then per-prompt rounds forced" >> $O/r42_ab.txt
  timeout -s KILL 300 python scripts/ab_step.py 10000000 $n KVIDX_CLASSES_MIN=1 KVIDX_CLASSES_MIN=100000000 >> $O/r42_ab.txt 2>&1
done
cat $O/r42_ab.txt

#!/bin/bash
# A/B: staging depth / CTA size of the class pipeline's kernel H, with kernel G on low-priority streams (the default now)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r27_ab.txt
echo "== 2 stages x 256 threads (default)" >> $O/r27_ab.txt
timeout -s KILL 300 python scripts/ab_step.py 10000000 1048576 base >> $O/r27_ab.txt 2>&1
for v in s3t128 s3t64 s4t64 s2t64; do
  echo "== $v" >> $O/r27_ab.txt
  KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/$v/libkvidx.so timeout -s KILL 300 python scripts/ab_step.py 10000000 1048576 base >> $O/r27_ab.txt 2>&1
done
cat $O/r27_ab.txt

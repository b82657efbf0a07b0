#!/bin/bash
# A/B: L2 evict_first for the token stream / dense rows (lib) vs without (lib_exp/nol2); kernel H prefetching its chunks to L2
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
C="base KVIDX_GROUP_SERIAL=2 KVIDX_HASH_PREFETCH=1 KVIDX_HASH_PREFETCH=2 KVIDX_GROUP_SERIAL=2,KVIDX_HASH_PREFETCH=1 KVIDX_GROUP_SERIAL=2,KVIDX_HASH_PREFETCH=2 KVIDX_GROUP_SERIAL=1,KVIDX_ROUNDS_PARTS=2,KVIDX_GROUP_SERIAL_GRID=3,KVIDX_HASH_PREFETCH=1"
echo "== L2 stream hints" > $O/r16_ab.txt
timeout -s KILL 600 python scripts/ab_step.py 10000000 1048576 $C >> $O/r16_ab.txt 2>&1
echo "== no hints" >> $O/r16_ab.txt
KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/nol2/libkvidx.so timeout -s KILL 600 python scripts/ab_step.py 10000000 1048576 $C >> $O/r16_ab.txt 2>&1
cat $O/r16_ab.txt
KVIDX_GROUP_SERIAL=2 KVIDX_HASH_PREFETCH=1 timeout -s KILL 300 python scripts/timeline.py 10000000 1048576 $O/r16_tl.json > $O/r16_tl.out 2>&1; tail -2 $O/r16_tl.out

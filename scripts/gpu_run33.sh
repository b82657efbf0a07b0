#!/bin/bash
# 2 GPUs: peer probes fetching the slot pair at once vs home slot first (one process holding both shards)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r33_pair.txt
for pp in 0 1 0 1; do
  echo "KVIDX_PEER_PAIR=$pp" >> $O/r33_pair.txt
  KVIDX_PEER_PAIR=$pp timeout -s KILL 300 python scripts/prof_sharded.py 10000000 1048576 2 >> $O/r33_pair.txt 2>&1
done
KVIDX_PEER_PAIR=0 timeout -s KILL 300 python scripts/prof_sharded.py 10000000 1048576 1 >> $O/r33_pair.txt 2>&1
cat $O/r33_pair.txt

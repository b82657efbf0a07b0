#!/bin/bash
# warp-per-prompt kernel P + 3-stage kernel H of the per-prompt rounds: parity on every path, then step times with / without
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r22_ab.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q > $O/r22_pytest.log 2>&1; tail -3 $O/r22_pytest.log
for n in 32768 65536 131072 262144; do
  echo "== $n prompts" >> $O/r22_ab.txt
  timeout -s KILL 300 python scripts/ab_step.py 10000000 $n KVIDX_ROUNDS_WARP_MAX=1000000 KVIDX_ROUNDS_WARP_MAX=1000000,KVIDX_ROUNDS_WARP_STAGES=2 KVIDX_ROUNDS_WARP=0 >> $O/r22_ab.txt 2>&1
done
cat $O/r22_ab.txt
KVIDX_ROUNDS_WARP_MAX=1000000 timeout -s KILL 300 python scripts/timeline.py 10000000 65536 $O/r22_tl64k.json > $O/r22_tl.out 2>&1; tail -1 $O/r22_tl.out

#!/bin/bash
# warp-per-prompt kernel P of the per-prompt rounds: parity on every path, then step times at 32 Ki .. 256 Ki prompts with and without it
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q > $O/r21_pytest.log 2>&1; tail -3 $O/r21_pytest.log
for n in 32768 65536 131072 262144; do
  echo "== $n prompts" >> $O/r21_ab.txt
  timeout -s KILL 300 python scripts/ab_step.py 10000000 $n KVIDX_ROUNDS_WARP_MAX=1000000 KVIDX_ROUNDS_WARP=0 KVIDX_ROUNDS_WARP_MAX=1000000,KVIDX_ROUNDS_OVERLAP=0 >> $O/r21_ab.txt 2>&1
done
cat $O/r21_ab.txt
KVIDX_ROUNDS_WARP_MAX=1000000 timeout -s KILL 300 python scripts/timeline.py 10000000 65536 $O/r21_tl64k.json > $O/r21_tl.out 2>&1; tail -1 $O/r21_tl.out

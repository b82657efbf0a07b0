#!/bin/bash
# walk kernel of the warp-per-prompt rounds: survivors appended per CTA iteration (two barriers) vs one atomic per surviving warp
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out; rm -f $O/r37_ab.txt
KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/wa/libkvidx.so timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rounds" > $O/r37_pytest.log 2>&1; tail -2 $O/r37_pytest.log
timeout -s KILL 300 python -m pytest tests/test_service.py -m gpu -x -q > $O/r37_pytest_service.log 2>&1; tail -2 $O/r37_pytest_service.log
for n in 8192 16384 32768 49152; do
  echo "== $n prompts: CTA-aggregated append (default), then per-warp atomics" >> $O/r37_ab.txt
  timeout -s KILL 300 python scripts/ab_step.py 10000000 $n base >> $O/r37_ab.txt 2>&1
  KVIDX_LIB=$PWD/llm-d-kv-cache-manager_b200/lib_exp/wa/libkvidx.so timeout -s KILL 300 python scripts/ab_step.py 10000000 $n base >> $O/r37_ab.txt 2>&1
done
cat $O/r37_ab.txt

#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_coop2 scripts/ubench_coop2.cu > $O/r5_ubench2.log 2>&1
timeout -s KILL 60 scripts/ubench_coop2 >> $O/r5_ubench2.log 2>&1
cat $O/r5_ubench2.log | tail -8
timeout -s KILL 120 llm-d-kv-cache-manager_b200/lib/kvidx_qps 1000 2.0 4096 4096 > $O/r5_qps1000.json 2>&1; cat $O/r5_qps1000.json
timeout -s KILL 120 llm-d-kv-cache-manager_b200/lib/kvidx_qps 64 2.0 4096 4096 > $O/r5_qps64.json 2>&1; cat $O/r5_qps64.json
timeout -s KILL 120 llm-d-kv-cache-manager_b200/lib/kvidx_qps 1 2.0 4096 4096 > $O/r5_qps1.json 2>&1; cat $O/r5_qps1.json
KVIDX_SUBMIT_QUEUE=0 timeout -s KILL 120 llm-d-kv-cache-manager_b200/lib/kvidx_qps 1000 2.0 4096 4096 > $O/r5_qps1000_noqueue.json 2>&1; cat $O/r5_qps1000_noqueue.json
timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $O/r5_tests.log 2>&1; tail -4 $O/r5_tests.log

#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_coop2 scripts/ubench_coop2.cu > $O/r6_ubench2.log 2>&1
timeout -s KILL 60 scripts/ubench_coop2 >> $O/r6_ubench2.log 2>&1
cat $O/r6_ubench2.log | tail -9
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_coop scripts/ubench_coop.cu > $O/r6_ubench.log 2>&1
timeout -s KILL 60 scripts/ubench_coop >> $O/r6_ubench.log 2>&1; tail -2 $O/r6_ubench.log
timeout -s KILL 200 python scripts/lat.py coop > $O/r6_lat_coop.log 2>&1; cat $O/r6_lat_coop.log | cut -c1-800
for t in 1000 64 8; do timeout -s KILL 120 llm-d-kv-cache-manager_b200/lib/kvidx_qps $t 2.0 4096 4096 > $O/r6_qps$t.json 2>&1; cat $O/r6_qps$t.json; done
timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $O/r6_tests.log 2>&1; tail -4 $O/r6_tests.log

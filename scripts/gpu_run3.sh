#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_coop scripts/ubench_coop.cu > $O/r3_ubench.log 2>&1
timeout -s KILL 60 scripts/ubench_coop >> $O/r3_ubench.log 2>&1
timeout -s KILL 300 python scripts/lat.py all > $O/r3_lat.log 2>&1
KVIDX_ZEROCOPY_MAX=0 timeout -s KILL 200 python scripts/lat.py coop > $O/r3_lat_coop_nozc.log 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $O/r3_tests.log 2>&1
tail -6 $O/r3_tests.log; cat $O/r3_ubench.log | tail -4; cat $O/r3_lat.log

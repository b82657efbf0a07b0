#!/bin/bash
# second GPU session of round 2: microbench, latency, tests, bench, TMA A/B, ncu captures.  Everything lands in gpurun_out/.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_coop scripts/ubench_coop.cu > $O/r2_ubench.log 2>&1
timeout -s KILL 60 scripts/ubench_coop >> $O/r2_ubench.log 2>&1
if ! timeout -s KILL 200 python scripts/lat.py coop > $O/r2_lat_coop.log 2>&1; then
  echo "zero-copy path failed; disabling" >> $O/r2_lat_coop.log
  export KVIDX_ZEROCOPY_MAX=0
  timeout -s KILL 200 python scripts/lat.py coop >> $O/r2_lat_coop.log 2>&1
fi
KVIDX_ZEROCOPY_MAX=0 timeout -s KILL 200 python scripts/lat.py coop > $O/r2_lat_coop_nozc.log 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider > $O/r2_tests.log 2>&1
tail -4 $O/r2_tests.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
tail -c 600 $O/r2_bench_n1.json
KVIDX_GROUP_TMA=0 KVIDX_BENCH_QUICK=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > $O/r2_bench_tma0.json 2> $O/r2_bench_tma0.err
KVIDX_GROUP_TMA=1 KVIDX_BENCH_QUICK=1 timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > $O/r2_bench_tma1.json 2> $O/r2_bench_tma1.err
cat $O/r2_bench_tma0.json $O/r2_bench_tma1.json
# launch list of the default step with DRAM bytes (time shares + traffic)
KVIDX_BENCH_QUICK=1 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 > $O/r2_ncu_launches.out 2>&1
# full captures: token-streaming kernel with TMA and with cp.async, the cooperative small-batch kernel, the write path
KVIDX_BENCH_QUICK=1 KVIDX_BENCH_BATCH=524288 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:group_round_kernel -s 16 -c 2 -o $O/r2_group_round_tma python bench.py --steps 1 --warmup 1 > $O/r2_ncu_g1.out 2>&1
KVIDX_GROUP_TMA=0 KVIDX_BENCH_QUICK=1 KVIDX_BENCH_BATCH=524288 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:group_round_kernel -s 16 -c 2 -o $O/r2_group_round_cpasync python bench.py --steps 1 --warmup 1 > $O/r2_ncu_g0.out 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:coop_score_kernel -s 6 -c 1 -o $O/r2_coop python scripts/lat.py coop 65536 > $O/r2_ncu_coop.out 2>&1
KVIDX_BENCH_QUICK=1 KVIDX_BENCH_BATCH=65536 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"hash_events_kernel|apply_events_kernel" -s 4 -c 2 -o $O/r2_write python bench.py --steps 1 --warmup 1 > $O/r2_ncu_w.out 2>&1
ls -la $O | grep r2_

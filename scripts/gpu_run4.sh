#!/bin/bash
# 2-GPU session: sharded tests on real peers, bench at N=2 (sharded + replicas + routed all-to-all), NVLink counters of a sharded step.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi topo -m > $O/r4_topo.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_coop scripts/ubench_coop.cu > $O/r4_ubench.log 2>&1
timeout -s KILL 60 scripts/ubench_coop >> $O/r4_ubench.log 2>&1
tail -2 $O/r4_ubench.log
timeout -s KILL 200 python scripts/lat.py coop > $O/r4_lat_coop.log 2>&1
cat $O/r4_lat_coop.log | cut -c1-600
timeout -s KILL 900 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider --deselect tests/test_gpu_sharded.py --deselect tests/test_dist_gloo.py > $O/r4_tests_all.log 2>&1
tail -4 $O/r4_tests_all.log
timeout -s KILL 900 python -m pytest tests/test_gpu_sharded.py tests/test_dist_gloo.py -m gpu -q -p no:cacheprovider > $O/r4_tests.log 2>&1
grep -A25 'failed:' $O/r4_tests.log | head -40
tail -5 $O/r4_tests.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r4_bench_n2.json 2> $O/r4_bench_n2.err
tail -c 1500 $O/r4_bench_n2.json; tail -5 $O/r4_bench_n2.err
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlink" | head -40 > $O/r4_nvl_metrics.txt
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_aperture_peer.sum,lts__t_sectors_srcunit_tex_aperture_peer_lookup_miss.sum --clock-control none -k regex:"round|rounds_init|Radix|count_distinct" -c 1200 --csv --log-file $O/r4_sharded_launches.csv python scripts/prof_sharded.py 2097152 524288 2 > $O/r4_prof_sharded.out 2>&1
tail -3 $O/r4_prof_sharded.out
timeout -s KILL 300 python scripts/prof_sharded.py 2097152 524288 2 > $O/r4_sharded_plain.out 2>&1
cat $O/r4_sharded_plain.out | tail -2
ls -la $O | grep r4_

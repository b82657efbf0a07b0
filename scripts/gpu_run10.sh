#!/bin/bash
# 2-GPU session: where does a sharded step lose time?  per-kernel durations sharded vs unsharded (same script), peer read rate
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_peer scripts/ubench_peer.cu > $O/r10_peer.log 2>&1
timeout -s KILL 120 scripts/ubench_peer >> $O/r10_peer.log 2>&1; cat $O/r10_peer.log
for w in 1 2; do
  timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"round|group_lists|rounds_init|Radix|count_distinct" -c 1200 --csv --log-file $O/r10_launches_w$w.csv python scripts/prof_sharded.py 4194304 1048576 $w > $O/r10_prof_w$w.out 2>&1
  timeout -s KILL 200 python scripts/prof_sharded.py 4194304 1048576 $w > $O/r10_plain_w$w.out 2>&1; tail -1 $O/r10_plain_w$w.out
done

#!/bin/bash
# final single-GPU captures of round 2: bench, reference arm, complete step launch list (time + DRAM bytes), kernel profiles
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > $O/r7_bench_n1.json 2> $O/r7_bench_n1.err
tail -c 400 $O/r7_bench_n1.json; tail -3 $O/r7_bench_n1.err
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r7_bench_ref.json 2> $O/r7_bench_ref.err
tail -c 300 $O/r7_bench_ref.json
KVIDX_BENCH_QUICK=1 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"round|Radix|count_distinct|events|parents|offer_blocks" -c 2600 --csv --log-file $O/r7_launches.csv python bench.py --steps 1 --warmup 1 > $O/r7_ncu_launches.out 2>&1
tail -2 $O/r7_ncu_launches.out | cut -c1-300
# token-streaming kernel over the WHOLE batch as one part, cp.async vs TMA (DRAM-active comparison)
KVIDX_ROUNDS_PARTS=1 KVIDX_GROUP_TMA=0 KVIDX_BENCH_QUICK=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:group_round_kernel -s 1 -c 1 -o $O/r7_group_whole_cpasync python bench.py --steps 1 --warmup 1 > $O/r7_ncu_gw0.out 2>&1
KVIDX_ROUNDS_PARTS=1 KVIDX_GROUP_TMA=1 KVIDX_BENCH_QUICK=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:group_round_kernel -s 1 -c 1 -o $O/r7_group_whole_tma python bench.py --steps 1 --warmup 1 > $O/r7_ncu_gw1.out 2>&1
# cooperative kernel with 1024 prompts in flight (enough samples for the source view)
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:coop_score_kernel -s 220 -c 1 -o $O/r7_coop1024 python scripts/lat.py coop > $O/r7_ncu_coop.out 2>&1
# write path kernels of the fill
KVIDX_BENCH_QUICK=1 KVIDX_BENCH_BATCH=65536 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"hash_events_kernel|apply_events_kernel" -s 4 -c 2 -o $O/r7_write python bench.py --steps 1 --warmup 1 > $O/r7_ncu_w.out 2>&1
# speculative rounds A/B at 64K
for sp in 0 1; do KVIDX_ROUNDS_SPEC=$sp KVIDX_BENCH_QUICK=1 KVIDX_BENCH_BATCH=65536 timeout -s KILL 300 python bench.py --steps 20 --warmup 3 > $O/r7_bench_64k_spec$sp.json 2>/dev/null; cat $O/r7_bench_64k_spec$sp.json | cut -c1-200; done
ls -la $O | grep r7_

#!/bin/bash
# final validation of the round's code + ncu --set full of the two new kernels of the medium-batch path
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/r36_pytest.log 2>&1; tail -3 $O/r36_pytest.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:probe_round_warp_kernel -s 16 -c 1 -f -o $O/r36_probe_warp python scripts/ab_step.py 10000000 32768 base > $O/r36_ncu1.out 2>&1; tail -1 $O/r36_ncu1.out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"plain.*hash_round_kernel" -s 16 -c 1 -f -o $O/r36_hash3 python scripts/ab_step.py 10000000 32768 base > $O/r36_ncu2.out 2>&1; tail -1 $O/r36_ncu2.out
python scripts/ncu_kernel_summary.py $O/r36_probe_warp.ncu-rep $O/r36_probe_warp_kernel.json "plain::probe_round_warp_kernel (warp per prompt), round 0 of one half of a 32 768-prompt step (16 384 prompts), ncu --set full" 2>&1 | tail -1
python scripts/ncu_kernel_summary.py $O/r36_hash3.ncu-rep $O/r36_hash3_kernel.json "plain::hash_round_kernel<16, prompt-major keys, 3 stages>, 128-thread CTAs, round 0 of one half of a 32 768-prompt step, ncu --set full" 2>&1 | tail -1
timeout -s KILL 600 python bench.py > $O/r36_bench.json 2> $O/r36_bench.err; tail -c 300 $O/r36_bench.json

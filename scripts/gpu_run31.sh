#!/bin/bash
# 4-GPU bench (sharded index = value, replicas beside it)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r31_bench_n4.json 2> $O/r31_bench_n4.err
tail -c 400 $O/r31_bench_n4.json; tail -3 $O/r31_bench_n4.err

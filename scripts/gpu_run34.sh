#!/bin/bash
# 8-GPU bench on the round's final code (sharded index = value, replicas, all-to-all comparison, config #4)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r34_bench_n8.json 2> $O/r34_bench_n8.err
tail -c 400 $O/r34_bench_n8.json; tail -3 $O/r34_bench_n8.err

#!/bin/bash
# 2-GPU bench with the round's final defaults (sharded index = value, replicas beside it) + the multi-GPU tests
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $O/r30_pytest.log 2>&1; tail -2 $O/r30_pytest.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r30_bench_n2.json 2> $O/r30_bench_n2.err
tail -c 600 $O/r30_bench_n2.json; tail -3 $O/r30_bench_n2.err

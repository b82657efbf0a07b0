#!/bin/bash
# validation of the new defaults (kernel G on low-priority streams, L2 evict_first for streams): GPU suite + default bench
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > $O/r19_pytest.log 2>&1; tail -3 $O/r19_pytest.log
timeout -s KILL 600 python bench.py > $O/r19_bench.json 2> $O/r19_bench.err; tail -c 600 $O/r19_bench.json
timeout -s KILL 300 python scripts/ab_step.py 10000000 524288 base KVIDX_GROUP_SERIAL=0 > $O/r19_ab512k.txt 2>&1; cat $O/r19_ab512k.txt
timeout -s KILL 300 python scripts/ab_step.py 10000000 65536 base > $O/r19_ab64k.txt 2>&1; cat $O/r19_ab64k.txt

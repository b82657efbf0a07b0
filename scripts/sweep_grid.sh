# experiment helper: run the bench with a few settings of the class pipeline (one line each)
export KVIDX_BENCH_SKIP_CPU=1 KVIDX_BENCH_NOCHECK=1
run() { python bench.py --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3))"; }
for st in -1 0 1 2; do for pp in 4 8 12; do KVIDX_ROUNDS_STAGGER=$st KVIDX_ROUNDS_PARTS=$pp run stagger${st}_parts$pp; done; done

# timing ablations of the score kernel (results are WRONG in ablated builds; timing only)
for a in 1 4 8 13; do
KVIDX_LIB=$PWD/abl/abl_$a/libkvidx.so KVIDX_BENCH_NOCHECK=1 KVIDX_BENCH_SKIP_CPU=1 KVIDX_BENCH_E2E_BATCH=4096 timeout 600 python bench.py --steps 5 --warmup 3 2>/tmp/err_$a.log > gpurun_out/abl_$a.json
python -c "
import json; d=json.load(open('gpurun_out/abl_$a.json')); print('ablate $a', d['value'], d['ms_per_step'])" || tail -3 /tmp/err_$a.log
done

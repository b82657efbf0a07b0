for c in 2 3 4 5; do
KVIDX_SCORE_CTAS_PER_SM=$c KVIDX_BENCH_SKIP_CPU=1 KVIDX_BENCH_E2E_BATCH=4096 timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null > gpurun_out/occ_$c.json
python -c "
import json; d=json.load(open('gpurun_out/occ_$c.json')); print('ctas/sm $c', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done

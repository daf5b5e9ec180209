timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_plan_parity.py tests/test_gpu_full_size.py tests/test_gpu_leaf_search.py -x -q 2>&1 | tail -8
tools/ab_variants.sh v3 v4 v3 v4 > gpurun_out/ab_v4.log 2>&1; cat gpurun_out/ab_v4.log
timeout 400 python tools/bench_configs.py --only C2,X2d,X2s,X2t,C1b --reps 10 --no-oracle --out gpurun_out/cfg_x4.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:50], 'dev_us %.1f main %.1f postings %d'%(d['device_us'],d['k_window_collect_us'],d['postings_scored']))
"

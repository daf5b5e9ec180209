timeout 400 python tools/bench_configs.py --only C2,X2d,X2s,X2t --reps 10 --no-oracle --out gpurun_out/cfg_x2.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:50], 'dev_us %.1f main %.1f postings %d'%(d['device_us'],d['k_window_collect_us'],d['postings_scored']))
"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2d.log 2>&1; tail -3 gpurun_out/t_r2d.log

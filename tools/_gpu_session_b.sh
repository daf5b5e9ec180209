mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/b_pytest.log
tools/ab_variants.sh base_old new base_old new 2>&1 | tee gpurun_out/b_ab.log
for h in 1 2 8; do echo "HDIV=$h"; QWGPU_HDIV=$h tools/ab_variants.sh new 2>&1 | tee -a gpurun_out/b_ab.log; done
timeout 300 python tools/bench_configs.py --only C1,C1b,C2,C3,C3b --reps 10 --no-oracle --out gpurun_out/b_cfg.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:50], 'dev_us %.1f main %.1f wall %.1f'%(d['device_us'],d['k_window_collect_us'],d['leaf_search_wall_us']))
"

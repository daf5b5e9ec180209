mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/h_pytest.log
timeout 600 python bench.py > gpurun_out/h_bench_1gpu.json 2> gpurun_out/h_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/h_bench_1gpu.err; python tools/bench_line.py gpurun_out/h_bench_1gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/h_bench_1gpu.json'))
for k,v in d['configs'].items(): print(k, 'device_us %.1f main %.1f p50 %.3f frac %.4f'%(v['device_us'], v['main_kernel_us'], v['leaf_search_p50_ms'], v['roofline']['frac']))
c5=d['config5_mixed']; print(json.dumps({k:c5[k] for k in ('qps','latency_ms','hbm','device_us_alone_by_type','mean_latency_ms_by_type')}, indent=1))
PY

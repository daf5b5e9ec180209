mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/final_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_union -c 3 -o gpurun_out/final_k_union -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/final_ncu_union.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_select|k_pick" -c 4 -o gpurun_out/final_tail -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/final_ncu_tail.log 2>&1
timeout 600 python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/final_bench_1gpu.err; python tools/bench_line.py gpurun_out/final_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench_1gpu.json'))
for k,v in d['configs'].items(): print(k, 'device_us %.1f main %.1f p50 %.3f frac %.4f'%(v['device_us'], v['main_kernel_us'], v['leaf_search_p50_ms'], v['roofline']['frac']))
c5=d['config5_mixed']; print(json.dumps({k:c5[k] for k in ('qps','latency_ms','hbm')}))
print(d.get('cpu_baseline')); print(d['e2e']['single_query_latency_ms'], d['clocks'])
r=json.load(open('gpurun_out/final_bench_reference_arm.json')); print('reference arm value %.3f G'%(r['value']/1e9))
PY

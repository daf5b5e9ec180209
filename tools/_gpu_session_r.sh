mkdir -p gpurun_out
timeout 100 python bench.py --splits 4 --docs-per-split 200000 --steps 5 --no-cpu-baseline > gpurun_out/r_bench_small.json 2> gpurun_out/r_bench_small.err; echo "bench rc=$?"; tail -2 gpurun_out/r_bench_small.err
python -c "
import json; d=json.load(open('gpurun_out/r_bench_small.json')); print('value %.2fG'%(d['value']/1e9), list(d['configs'].keys()) if isinstance(d.get('configs'),dict) else d.get('configs')); c5=d['config5_mixed']; print(c5.get('error') or (c5['qps'], c5['latency_ms']['p50']))"

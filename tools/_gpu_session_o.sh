mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 > gpurun_out/o_bench_2gpu.json 2> gpurun_out/o_bench_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/o_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/o_bench_2gpu.json'))
print('value %.1fG e2e %.1fG ms/step %.3f frac %.4f'%(d['value']/1e9, d['e2e']['value']/1e9, d['ms_per_step'], d['roofline']['frac']))
print(json.dumps(d['e2e']['phase_ms_per_step']), d['e2e'].get('exchange'))
print(json.dumps(d.get('config4_strong'))[:400])
c5=d.get('config5_mixed'); print(json.dumps({k:c5[k] for k in ('qps','latency_ms','aggregate','hbm')} if c5 else None))
PY

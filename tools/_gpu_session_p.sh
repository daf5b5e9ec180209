mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python tools/bench_c5.py --concurrency 1,64 --out gpurun_out/p_c5.json 2>&1 | tail -2 | tee gpurun_out/p_c5.log
python - <<'PY'
import json
r=json.load(open('gpurun_out/p_c5.json'))[-1]
print(json.dumps(r['device_us_alone_by_type']))
PY

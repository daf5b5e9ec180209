mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d_pytest.log
QWGPU_STATIC_WORK=1 timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
tools/ab_variants.sh base_old new 2>&1 | tee gpurun_out/d_ab.log
echo STATIC; QWGPU_STATIC_WORK=1 tools/ab_variants.sh new 2>&1 | tee -a gpurun_out/d_ab.log
tools/ab_variants.sh new 2>&1 | tee -a gpurun_out/d_ab.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/d_launches.csv")) if len(r)>10]
hdr=rows[0]; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    d=dict(zip(hdr,r)); k=d["Kernel Name"][:40]; agg[k][0]+=1; agg[k][1]+=float(d["Metric Value"])/1000
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{k:42s} n={n:4d} avg {t/n:8.2f} us")
PY

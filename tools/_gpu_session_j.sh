mkdir -p gpurun_out
for uw in 15360 18432 20480 23552; do echo "UW=$uw"; QWGPU_UW=$uw tools/ab_variants.sh base 2>&1 | tee -a gpurun_out/j_ab.log; done
QWGPU_TRACE=1 timeout 300 python tools/bench_c5.py --concurrency 1 --announce --out gpurun_out/j_c5.json 2> gpurun_out/j_trace.log | tail -1
python - <<'PY'
import re, collections
cur=None; agg=collections.defaultdict(list)
for l in open('gpurun_out/j_trace.log'):
    m=re.match(r"\[c5\] (\S+)", l)
    if m: cur=m.group(1); continue
    m=re.search(r"leaf_search: decode (\d+) us, compile\+search (\d+) us \(engine wall (\d+) us, device (\d+) us, (\d+) launches\), merge (\d+) us, encode (\d+) us", l)
    if m and cur: agg[cur].append([int(x) for x in m.groups()])
for k,v in agg.items():
    v=v[-8:]; n=len(v); mean=[sum(x[i] for x in v)/n for i in range(7)]
    print(f"{k:28s} decode {mean[0]:6.0f} compile+search {mean[1]:7.0f} (engine wall {mean[2]:7.0f}, device {mean[3]:7.0f}, launches {mean[4]:3.0f}) merge {mean[5]:6.0f} encode {mean[6]:5.0f}")
PY
grep "search:" gpurun_out/j_trace.log | tail -12

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_pytest.log
for sc in 16 23 32 16; do echo "STRIDE_CAP=$sc"; QWGPU_STRIDE_CAP=$sc tools/ab_variants.sh base 2>&1 | tee -a gpurun_out/f_ab.log; done
timeout 600 python bench.py > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench_1gpu.err; python tools/bench_line.py gpurun_out/f_bench_1gpu.json
python -c "
import json; d=json.load(open('gpurun_out/f_bench_1gpu.json')); print(json.dumps(d.get('config5_mixed'), indent=1)[:3000]); print(d['setup'])"

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
tools/ab_variants.sh nodefer base nodefer base 2>&1 | tee gpurun_out/l_ab.log
timeout 400 python tools/bench_c5.py --concurrency 16,64 --out gpurun_out/l_c5.json 2>&1 | tail -2 | tee gpurun_out/l_c5.log
QWGPU_MAX_IN_FLIGHT=8 timeout 400 python tools/bench_c5.py --concurrency 64 --out gpurun_out/l_c5_8.json 2>&1 | tail -1 | tee -a gpurun_out/l_c5.log
QWGPU_MAX_IN_FLIGHT=32 timeout 400 python tools/bench_c5.py --concurrency 64 --out gpurun_out/l_c5_32.json 2>&1 | tail -1 | tee -a gpurun_out/l_c5.log

QWGPU_LIB=$PWD/quickwit_b200/libqwgpu_prof.so QWGPU_UPROF=1 timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs 2>&1 | grep -E "uprof|Error|error" | tail -3 > gpurun_out/uprof_ord.log; cat gpurun_out/uprof_ord.log
tools/ab_variants.sh base v2 base v2 > gpurun_out/ab_v2.log 2>&1; cat gpurun_out/ab_v2.log
timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_plan_parity.py -x -q 2>&1 | tail -3

QWGPU_FREE_UNION=1 timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_goldens.py -x -q > gpurun_out/t_fuzz.log 2>&1; tail -15 gpurun_out/t_fuzz.log
i=0
for v in "QWGPU_FREE_UNION=1" "A=1" "QWGPU_FREE_UNION=1 QWGPU_UW=5120" "QWGPU_FREE_UNION=1 QWGPU_UW=7168"; do
  i=$((i+1)); echo "== $v"; env $v timeout 200 python bench.py --steps 10 --no-cpu-baseline --no-configs > gpurun_out/ab_free_$i.json 2>/dev/null; python tools/bench_line.py gpurun_out/ab_free_$i.json
done > gpurun_out/ab_free.log 2>&1; cat gpurun_out/ab_free.log
QWGPU_FREE_UNION=1 QWGPU_LIB=$PWD/quickwit_b200/libqwgpu_prof.so QWGPU_UPROF=1 timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-configs 2>&1 | grep uprof | tail -2 > gpurun_out/uprof_free.log; cat gpurun_out/uprof_free.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2b.log 2>&1; tail -5 gpurun_out/t_r2b.log
QWGPU_FREE_UNION=1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2b_free.log 2>&1; tail -5 gpurun_out/t_r2b_free.log
QWGPU_DRIVER=1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2b_drv.log 2>&1; tail -5 gpurun_out/t_r2b_drv.log
timeout 400 python tools/bench_configs.py --reps 10 --no-oracle --out gpurun_out/cfg_r2b.json > gpurun_out/cfg_r2b.log 2>&1; tail -3 gpurun_out/cfg_r2b.log
QWGPU_DRIVER=1 timeout 400 python tools/bench_configs.py --reps 10 --no-oracle --out gpurun_out/cfg_r2b_drv.json > gpurun_out/cfg_r2b_drv.log 2>&1; tail -3 gpurun_out/cfg_r2b_drv.log

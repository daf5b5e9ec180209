QWGPU_DRIVER=1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2c_drv.log 2>&1; tail -5 gpurun_out/t_r2c_drv.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2c.log 2>&1; tail -5 gpurun_out/t_r2c.log
QWGPU_DRIVER=1 timeout 400 python tools/bench_configs.py --reps 10 --no-oracle --out gpurun_out/cfg_r2c_drv.json > gpurun_out/cfg_r2c_drv.log 2>&1; tail -8 gpurun_out/cfg_r2c_drv.log

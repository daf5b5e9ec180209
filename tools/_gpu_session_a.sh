mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
bash tools/_gpu_prof.sh

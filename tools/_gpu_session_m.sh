mkdir -p gpurun_out
tools/ab_variants.sh base t384 t448 base t384 2>&1 | tee gpurun_out/m_ab.log
timeout 600 python -m pytest tests/test_gpu_leaf_search.py -m gpu -x -q 2>&1 | tail -3
QWGPU_LIB=$PWD/quickwit_b200/libqwgpu_t384.so timeout 600 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2

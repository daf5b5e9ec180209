mkdir -p gpurun_out
for sc in 16 23 32; do echo "STRIDE_CAP=$sc"; QWGPU_STRIDE_CAP=$sc tools/ab_variants.sh base 2>&1 | tee -a gpurun_out/e_ab.log; done
QWGPU_STRIDE_CAP=16 tools/ab_variants.sh base 2>&1 | tee -a gpurun_out/e_ab.log

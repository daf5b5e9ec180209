mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_select|k_merge|k_pick" -c 8 -o gpurun_out/r2_tail -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2_ncu_tail.log 2>&1
tools/ab_variants.sh base base

mkdir -p gpurun_out
nproc
QWGPU_UPROF=1 QWGPU_LIB=$PWD/quickwit_b200/libqwgpu_prof.so timeout 300 python tools/bench_configs.py --only C2 --reps 2 --no-oracle --out gpurun_out/i_c2.json 2>&1 | grep uprof | tail -3 | tee gpurun_out/i_uprof.log
echo "--- adaptive wait"; timeout 400 python tools/bench_c5.py --concurrency 1,8,64 --out gpurun_out/i_c5.json 2>&1 | tail -3 | tee gpurun_out/i_c5.log
echo "--- always spin"; QWGPU_SPIN_LIMIT=100000 timeout 400 python tools/bench_c5.py --concurrency 64 --out gpurun_out/i_c5_spin.json 2>&1 | tail -1 | tee -a gpurun_out/i_c5.log
echo "--- always block"; QWGPU_SPIN_LIMIT=0 timeout 400 python tools/bench_c5.py --concurrency 1,64 --out gpurun_out/i_c5_block.json 2>&1 | tail -2 | tee -a gpurun_out/i_c5.log

mkdir -p gpurun_out
# 1. launch list of the bench command (all kernels, durations only)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
# 2. full captures of the dominant kernels
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_union -c 3 -o gpurun_out/r2_k_union -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/r2_ncu_union.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_aggscan -c 1 -o gpurun_out/r2_k_aggscan -f python tools/bench_configs.py --only C4 --reps 1 --no-oracle --out gpurun_out/c4.json > gpurun_out/r2_ncu_c4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_driver -c 2 -o gpurun_out/r2_k_driver_c3 -f python tools/bench_configs.py --only C3 --reps 1 --no-oracle --out gpurun_out/c3.json > gpurun_out/r2_ncu_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_driver -c 2 -o gpurun_out/r2_k_driver_c1 -f python tools/bench_configs.py --only C1 --reps 1 --no-oracle --out gpurun_out/c1.json > gpurun_out/r2_ncu_c1.log 2>&1
# 3. launch lists of the single configs (every kernel of one query)
for c in C1 C3 C4; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_$c.csv python tools/bench_configs.py --only $c --reps 1 --no-oracle --out gpurun_out/x.json > /dev/null 2>&1
done
# 4. the bench line itself (not under a profiler)
timeout 400 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; python tools/bench_line.py gpurun_out/r2_bench_1gpu.json
ls -la gpurun_out/*.ncu-rep

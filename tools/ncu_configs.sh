#!/bin/bash
# ncu captures of the window engine on single configs (run on the GPU box through gpurun).
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:k_window"
$NCU -c 1 -o gpurun_out/prof_c4 -f python tools/bench_configs.py --only C4 --reps 1 --no-oracle --out gpurun_out/c4.json > gpurun_out/ncu_c4.log 2>&1
$NCU -c 3 -o gpurun_out/prof_c1 -f python tools/bench_configs.py --only C1 --reps 1 --no-oracle --out gpurun_out/c1.json > gpurun_out/ncu_c1.log 2>&1
$NCU -c 2 -o gpurun_out/prof_c3b -f python tools/bench_configs.py --only C3b --reps 1 --no-oracle --out gpurun_out/c3b.json > gpurun_out/ncu_c3b.log 2>&1
$NCU -c 2 -o gpurun_out/prof_c2 -f python tools/bench_configs.py --only C2 --reps 1 --no-oracle --out gpurun_out/c2.json > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep

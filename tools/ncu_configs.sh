#!/bin/bash
# ncu captures of the window engine on single configs (run on the GPU box through gpurun):
#   tools/ncu_configs.sh "C4:1 C3:3"   -> gpurun_out/prof_<cfg>.ncu-rep, <n> k_window launches each
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -k regex:k_window"
for spec in ${1:-C4:1}; do
  cfg=${spec%%:*}; n=${spec##*:}
  lc=$(echo $cfg | tr A-Z a-z)
  timeout 280 $NCU -c $n -o gpurun_out/prof_$lc -f python tools/bench_configs.py --only $cfg --reps 1 --no-oracle --out gpurun_out/$lc.json > gpurun_out/ncu_$lc.log 2>&1
done
ls -la gpurun_out/*.ncu-rep

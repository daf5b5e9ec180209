#!/bin/bash
# A/B of development builds of the library on the bench workload, back to back on one box:
#   make -C quickwit_b200/csrc QW_EXTRA="-DSOME_FLAG" && cp quickwit_b200/libqwgpu.so quickwit_b200/libqwgpu_X.so
#   tools/ab_variants.sh base X base X      ("base" = quickwit_b200/libqwgpu.so)
# bench.py picks the library through QWGPU_LIB (quickwit_b200/ffi.py).
for v in "$@"; do
  lib=$PWD/quickwit_b200/libqwgpu_$v.so
  [ "$v" = base ] && lib=$PWD/quickwit_b200/libqwgpu.so
  QWGPU_LIB=$lib timeout 120 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
b=json.loads(sys.stdin.read()); print('$v', 'value %.1f G' % (b['value']/1e9), 'e2e %.1f G' % (b['e2e']['value']/1e9), 'main %.1f us' % b['roofline']['avg_launch_us'], 'ms/step %.3f' % b['ms_per_step'])"
done

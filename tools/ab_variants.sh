#!/bin/bash
for v in "$@"; do
  QWGPU_LIB=$PWD/quickwit_b200/libqwgpu_$v.so timeout 120 python bench.py --steps 10 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
b=json.loads(sys.stdin.read()); print('$v', 'value %.1f G' % (b['value']/1e9), 'e2e %.1f G' % (b['e2e']['value']/1e9), 'main %.1f us' % b['roofline']['avg_launch_us'], 'ms/step %.3f' % b['ms_per_step'])"
done

#!/bin/bash
# steady-state throughput of several library builds: tools/gpu_variants.sh steps name1 name2 ...
STEPS=$1; shift
for v in "$@"; do
  lib=deepsignal_plant_b200/libdsp_b200_$v.so; [ "$v" = "main" ] && lib=deepsignal_plant_b200/libdsp_b200.so
  DSP_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', 'sites/s %.3fM' % (d['value']/1e6), 'e2e %.3fM' % (d['e2e']['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], d['roofline']['last_step_kernel_ms'], d['clocks'])"
done

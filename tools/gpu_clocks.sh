#!/bin/bash
# Long bench run with nvidia-smi sampling every 100 ms: shows steady-state clocks / power / throttle reasons.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu --format=csv,noheader -lms 100 > $OUT/clocks_$1.csv &
SMI=$!
sleep 1
timeout 600 python bench.py --steps ${2:-300} --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > $OUT/bench_long_$1.json
sleep 0.5
kill $SMI
python - <<PY
import re
rows=[l.strip().split(', ') for l in open('$OUT/clocks_$1.csv') if l.strip()]
print('samples',len(rows))
for r in rows[::max(1,len(rows)//40)]: print(r)
PY
cut -c1-400 $OUT/bench_long_$1.json

#!/bin/bash
# BASELINE.json configs[3]: throughput of the other model variants (steady state, 150 steps each)
for cfg in "both_bilstm 13 16" "seq_bilstm 13 16" "signal_bilstm 13 16" "both_bilstm 17 20"; do
  set -- $cfg
  timeout 400 python bench.py --steps 150 --warmup 3 --no-cpu-baseline --module $1 --seq_len $2 --signal_len $3 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1 bn$2_sn$3:', 'sites/s %.3fM' % (d['value']/1e6), 'e2e %.3fM' % (d['e2e']['value']/1e6), 'ms/step %.3f' % d['ms_per_step'], 'flop/site', r['flop_per_site'], 'whole_step_frac %.3f' % r['whole_step_frac'], r['last_step_kernel_ms'])"
done

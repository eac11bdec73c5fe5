#!/bin/bash
# launch list (ncu, serialised, cold) of the single-GPU call_freq aggregation + its throughput line
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 60 --csv \
    --log-file gpurun_out/launches_freq.csv python tools/bench_freq.py --cpu_sample 1000 > gpurun_out/ncu_freq_stdout.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_freq.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) != len(h):
        continue
    d = dict(zip(h, r))
    k = (d["ID"], d["Kernel Name"][:70])
    agg.setdefault(k, {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
seen = collections.OrderedDict()
for (i, k), m in agg.items():
    e = seen.setdefault(k, [0, 0.0, 0.0, 0.0])
    e[0] += 1; e[1] += m.get("gpu__time_duration.sum", 0); e[2] += m.get("dram__bytes_read.sum", 0); e[3] += m.get("dram__bytes_write.sum", 0)
for k, e in seen.items():
    print("%3d x %9.1f us  rd %8.1f MB  wr %8.1f MB  %s" % (e[0], e[1] / e[0] / 1e3, e[2] / e[0] / 1e6, e[3] / e[0] / 1e6, k))
PY

#!/bin/bash
# Run on the GPU box (via gpurun): tests, bench, launch list, full ncu capture of the layer kernels.
# Usage: tools/gpu_measure.sh <tag> [quick]
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_$TAG.json
if [ "$2" != "quick" ]; then
  # launch list: skip the 3 warm-up steps (9 launches each), list 2 timed steps + e2e
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 27 -c 27 --csv \
      --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_bench_$TAG.log 2>&1
  # full capture of one step's layer kernels (smaller batch keeps the replays short)
  timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:layer_kernel|branch_kernel" -s 24 -c 8 \
      -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --batch 37888 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
  ls -la $OUT | tail -8
fi

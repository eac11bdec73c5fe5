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
  # launch list (6 launches per step) + `ncu --set full` of every kernel of one step; summarise here with
  # python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep profiles/<name>.csv --traffic-json profiles/roofline_traffic.json
  bash tools/gpu_profile_r02.sh $TAG
  ls -la $OUT | tail -8
fi

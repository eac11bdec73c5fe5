#!/bin/bash
# Round 2 profile of one forward step (6 launches): launch list + `ncu --set full` of every kernel of one step.
# Usage (GPU box): bash tools/gpu_profile_r02.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
# launch list: skip the 3 warm-up steps (6 launches each), list the 2 timed steps and the e2e region
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 24 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-freq > $OUT/ncu_bench_$TAG.log 2>&1
# full capture of one step (smaller batch keeps the ~40 replays per kernel short)
timeout 1200 ncu --set full --clock-control none --import-source on -s 18 -c 6 \
    -o $OUT/prof_$TAG -f python bench.py --steps 1 --warmup 3 --batch 37888 --no-cpu-baseline --no-freq > $OUT/ncu_full_$TAG.log 2>&1
ls -la $OUT/prof_$TAG.ncu-rep $OUT/launches_$TAG.csv

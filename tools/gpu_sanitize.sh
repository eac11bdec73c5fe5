#!/bin/bash
# racecheck + synccheck of the tcgen05 kernels on a multi-wave grid; logs -> gpurun_out/
mkdir -p gpurun_out
SITES=${SITES:-40960}
for dual in ${DUALS:-0 1}; do
  for tool in synccheck racecheck; do
    echo "=== $tool dual=$dual sites=$SITES" 
    env $( [ "$dual" = fused ] || echo DSP_B200_BRANCH_DUAL=$dual ) timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --print-limit 20 \
        python tools/sanitize_run.py --sites $SITES > gpurun_out/sanitize_${tool}_dual${dual}.log 2>&1
    echo "rc=$?"; tail -4 gpurun_out/sanitize_${tool}_dual${dual}.log
  done
done

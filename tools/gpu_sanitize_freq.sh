#!/bin/bash
# racecheck + memcheck of the call_freq kernels, including the staged scatter and the exchange control kernels:
# 2 ranks sharing cuda:0 (CUDA-IPC windows, gloo control plane) under compute-sanitizer --target-processes all
mkdir -p gpurun_out
for tool in racecheck memcheck; do
  echo "=== $tool"
  timeout ${SAN_TIMEOUT:-500} compute-sanitizer --tool $tool --target-processes all --print-limit 10 \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2958${#tool} \
      tests/freq_dist_worker.py --tensor_check --records 600000 --prob_cf 0.3 > gpurun_out/sanitize_freq_${tool}.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|\"ok\"" gpurun_out/sanitize_freq_${tool}.log | tail -6
done

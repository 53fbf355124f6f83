#!/bin/bash
mkdir -p gpurun_out
for cfg in "5 0" "13 0" "5 1" "13 1"; do
  set -- $cfg
  echo "== MYOLO_FUSE_BN=$1 MYOLO_Y_OVERLAP=$2"
  for rep in 1 2; do
    MYOLO_FUSE_BN=$1 MYOLO_Y_OVERLAP=$2 timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu -k "replayed" 2>&1 | grep -E "AssertionError: \(|passed|failed" | cut -c1-200
  done
done 2>&1 | tee gpurun_out/r02h_replay_matrix.log

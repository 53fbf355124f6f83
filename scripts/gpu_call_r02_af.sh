#!/bin/bash
# FMA mask tail: which part of the epilogue is the time?  (MYOLO_WIN_BO experiment bits 1024 no weight loads, 2048 no dot products, 4096 no bias / ReLU)
mkdir -p gpurun_out
for bo in 0 1024 2048 6144 0; do
  MYOLO_WIN_BO=$bo timeout 120 python scripts/profile_conv.py 4704 20 deconv_h | tail -1 | sed "s/^/win_bo=$bo /"
done 2>&1 | tee gpurun_out/r02af_tail_parts.log

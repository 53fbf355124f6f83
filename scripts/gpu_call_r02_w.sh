#!/bin/bash
# what makes the fused BN-backward epilogue slow the main loop? standalone timing with parts of the epilogue switched off
mkdir -p gpurun_out
for bo in 0 128 256 384 0; do
  MYOLO_WIN_BO=$bo python scripts/profile_conv.py 4704 20 dgrad_h | tail -1 | sed "s/^/win_bo=$bo /"
done 2>&1 | tee gpurun_out/r02w_dgrad_parts.log
python scripts/profile_conv.py 4704 20 fwd_h | tail -1 | tee -a gpurun_out/r02w_dgrad_parts.log

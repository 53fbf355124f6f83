#!/bin/bash
# one `ncu --set full` capture per mask-head kernel (run alone at benchmark size) -> gpurun_out/ncu_<which>.{ncu-rep,txt}
# usage (under gpurun, one GPU): bash scripts/capture_ncu.sh fwd_h dgrad_h wgrad_h deconv_h
mkdir -p gpurun_out
for w in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:'tc_conv_win_kernel|tc_wgrad' -s 2 -c 1 -f \
      -o gpurun_out/ncu_$w python scripts/profile_conv.py 4704 2 $w > gpurun_out/ncu_$w.log 2>&1
  ncu -i gpurun_out/ncu_$w.ncu-rep --page raw --csv > gpurun_out/ncu_$w.raw.csv 2>/dev/null
  python scripts/profile_conv.py 4704 10 $w | tail -1
done

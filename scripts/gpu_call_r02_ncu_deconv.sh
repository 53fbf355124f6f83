#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'tc_conv_win_kernel' -s 2 -c 1 -f -o gpurun_out/ncu_deconv_h python scripts/profile_conv.py 4704 4 deconv_h > gpurun_out/ncu_deconv_h.log 2>&1
ncu -i gpurun_out/ncu_deconv_h.ncu-rep --page raw --csv > gpurun_out/ncu_deconv_h.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_deconv_h.ncu-rep --page source --csv > gpurun_out/ncu_deconv_h.source.csv 2>/dev/null
ls -la gpurun_out/ncu_deconv_h.*

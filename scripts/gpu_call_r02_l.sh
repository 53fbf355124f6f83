#!/bin/bash
mkdir -p gpurun_out
for v in alt_e alt_f alt_g alt_h; do
  if [ -n "$v" ]; then export MYOLO_LIB=$PWD/mask-yolo_b200/lib/$v.so; else unset MYOLO_LIB; fi
  echo "== variant ${v:-default(minblk3,ty8)}"
  python scripts/bench_hbm_kernels.py 2>&1 | grep -E "^dw_.*(dw1 |dw5 |total)"
done 2>&1 | tee gpurun_out/r02l_dw_variants.log

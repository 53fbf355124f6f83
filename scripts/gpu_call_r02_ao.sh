#!/bin/bash
# BN apply on load in the depthwise kernels (MYOLO_FUSE_BN bit 2) re-measured on the strip kernels of the final build
mkdir -p gpurun_out
for f in 29 31 29 31 29 31; do
  MYOLO_FUSE_BN=$f timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse_bn=$f', round(d['value'],1), d['ms_per_step'], d['gpu_launches']//d['steps'], 'launches')"
done | tee gpurun_out/r02ao_ab.log

#!/bin/bash
# stream switches re-measured on the final build: default | mask-head filter gradients on their own stream | no side stream for the backbone's filter gradients
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-24s' % '$1', round(d['value'],1), d['ms_per_step'])"
}
for i in 1 2 3; do
  run MYOLO_NOP=1
  run MYOLO_W_OVERLAP=1
  run MYOLO_BWD_STREAMS=0
done | tee gpurun_out/r02ap_ab.log

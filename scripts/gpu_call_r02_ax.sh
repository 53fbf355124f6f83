#!/bin/bash
# same-box check of the adopted stream schedule against the previous one (MYOLO_Y_SIDE=0 MYOLO_W_OVERLAP=0), alternating
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --config $2 --no-cpu-baseline --no-parity --no-fp32-class --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-3s %-40s' % ('$2', '$1'), round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'kernel', round(d['roofline']['avg_launch_ms'],4))"
}
OLD="MYOLO_Y_SIDE=0 MYOLO_W_OVERLAP=0"
for i in 1 2 3; do
  run "$OLD" c2
  run "MYOLO_NOP=1" c2
done | tee gpurun_out/r02ax_ab.log
for i in 1 2; do
  run "$OLD" c3
  run "MYOLO_NOP=1" c3
  run "$OLD" c5
  run "MYOLO_NOP=1" c5
done | tee -a gpurun_out/r02ax_ab.log

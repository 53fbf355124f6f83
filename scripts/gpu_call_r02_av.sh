#!/bin/bash
# the new stream schedule (yolo-branch filter gradients on their own stream; the last 3 mask-head filter gradients on theirs,
# sized for 110 SMs) at configs 3 and 5, and finer settings at config 2
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --config $2 --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-3s %-75s' % ('$2', '$1'), round(d['value'],1), round(d['ms_per_step'],3))"
}
NEW="MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=110"
for i in 1 2; do
  run "MYOLO_NOP=1" c3
  run "$NEW" c3
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=2 MYOLO_W_SMS=110" c3
  run "MYOLO_NOP=1" c5
  run "$NEW" c5
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=2 MYOLO_W_SMS=110" c5
  run "MYOLO_NOP=1" c2
  run "$NEW" c2
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=100" c2
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=120" c2
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=4 MYOLO_W_SMS=100" c2
done | tee gpurun_out/r02av_ab.log

#!/bin/bash
# A/B: only the LAST n mask-head filter gradients (MYOLO_W_DEFER) go to their own stream, sized for MYOLO_W_SMS SMs
mkdir -p gpurun_out
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-60s' % '$1', round(d['value'],1), round(d['ms_per_step'],3))"
}
for i in 1 2; do
  run "MYOLO_NOP=1"
  run "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=2 MYOLO_W_SMS=110"
  run "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=2 MYOLO_W_SMS=130"
  run "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=100"
  run "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=110"
  run "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=120"
  run "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=130"
  run "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=4 MYOLO_W_SMS=110"
done | tee gpurun_out/r02at_ab.log
bash scripts/phase_timeline.sh "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=110" "MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=2 MYOLO_W_SMS=110" 2>&1 | tee gpurun_out/r02at_phases.log

#!/bin/bash
# A/B: the yolo branch's filter gradients on their own stream (MYOLO_Y_SIDE), alone and with the deferred mask-head filter gradients
mkdir -p gpurun_out
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-75s' % '$1', round(d['value'],1), round(d['ms_per_step'],3))"
}
for i in 1 2; do
  run "MYOLO_NOP=1"
  run "MYOLO_Y_SIDE=1"
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=2 MYOLO_W_SMS=110"
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=110"
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=130"
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=4 MYOLO_W_SMS=110"
  run "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=5 MYOLO_W_SMS=110"
done | tee gpurun_out/r02au_ab.log
bash scripts/phase_timeline.sh "MYOLO_Y_SIDE=1" "MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=110" 2>&1 | tee gpurun_out/r02au_phases.log
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02au_tests_default.log
MYOLO_Y_SIDE=1 MYOLO_W_OVERLAP=1 MYOLO_W_DEFER=3 MYOLO_W_SMS=110 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02au_tests_switches.log

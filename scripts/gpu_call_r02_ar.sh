#!/bin/bash
# NOTE: MYOLO_WGRAD_SMS was this experiment's process-wide form of what became myolo_set_wgrad_sms / MYOLO_W_SMS
# (profiles/r02_wgrad_sm_subset_ab.txt); with the current tree use MYOLO_W_DEFER=5 MYOLO_W_SMS=<n> for the same runs.
# (1) GPU tests + timing of the two-launch polygon rasteriser; (2) A/B: mask-head filter gradients deferred to their own stream
# (MYOLO_W_OVERLAP=1) on a SUBSET of the SMs (MYOLO_WGRAD_SMS), so that the backbone's backward finds free SMs next to them
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_via_polygons.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02ar_polygon_tests.log
timeout 200 python scripts/polygon_raster_time.py 2>&1 | tail -2 | tee gpurun_out/r02ar_polygon_time.log
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-60s' % '$1', round(d['value'],1), round(d['ms_per_step'],3))"
}
for i in 1 2; do
  run "MYOLO_NOP=1"
  run "MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SMS=130"
  run "MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SMS=120"
  run "MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SMS=110"
  run "MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SMS=92"
  run "MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SMS=74"
done | tee gpurun_out/r02ar_ab.log

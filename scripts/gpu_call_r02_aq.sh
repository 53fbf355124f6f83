#!/bin/bash
# NOTE: MYOLO_BENCH_HIPRI / MYOLO_WGRAD_SPLIT_MUL / MYOLO_CARVEOUT were switches of this experiment only (all negative,
# profiles/r02_wgrad_overlap_priority_split_ab_negative.txt); they are not in the tree any more.
# (1) GPU tests of the VIA polygon rasteriser + its timing; (2) A/B: main chain on a high-priority stream, mask-head filter
# gradients deferred to their own stream (MYOLO_W_OVERLAP) as shorter CTAs (MYOLO_WGRAD_SPLIT_MUL), max-shared carve-out
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_via_polygons.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02aq_polygon_tests.log
timeout 200 python scripts/polygon_raster_time.py 2>&1 | tail -2 | tee gpurun_out/r02aq_polygon_time.log
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-90s' % '$1', round(d['value'],1), round(d['ms_per_step'],3))"
}
for i in 1 2; do
  run "MYOLO_NOP=1"
  run "MYOLO_BENCH_HIPRI=1"
  run "MYOLO_BENCH_HIPRI=1 MYOLO_W_OVERLAP=1"
  run "MYOLO_BENCH_HIPRI=1 MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SPLIT_MUL=2"
  run "MYOLO_BENCH_HIPRI=1 MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SPLIT_MUL=4"
  run "MYOLO_BENCH_HIPRI=1 MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SPLIT_MUL=4 MYOLO_CARVEOUT=1"
done | tee gpurun_out/r02aq_ab.log

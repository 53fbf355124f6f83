#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_h16_gpu.py -q -m gpu -x -k "bn" 2>&1 | tail -3
timeout 600 python scripts/flake_train_loop.py 12 2>&1 | grep "^run" | tee gpurun_out/r02ad_flake.log
OLD=$PWD/mask-yolo_b200/lib/alt_old.so
for t in old new old new; do
  if [ $t = old ]; then export MYOLO_LIB=$OLD; else unset MYOLO_LIB; fi
  timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse > gpurun_out/r02ad_bench_$t.json 2> gpurun_out/r02ad_bench_$t.err
  echo "lib=$t $(python - <<PY
import json
d=json.loads(open('gpurun_out/r02ad_bench_$t.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'])
PY
)" | tee -a gpurun_out/r02ad_ab.log
done

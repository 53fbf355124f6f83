#!/bin/bash
# ROIAlign forward with the feature columns of a sample row cached in registers: tests (bit-exact), kernel alone, step A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_h16_gpu.py -q -m gpu -x -k "roialign" 2>&1 | tail -3
OLD=$PWD/mask-yolo_b200/lib/alt_old.so
for t in old new old new; do
  if [ $t = old ]; then export MYOLO_LIB=$OLD; else unset MYOLO_LIB; fi
  python scripts/bench_hbm_kernels.py 2>&1 | grep -E "^roialign" | sed "s/^/lib=$t /"
done 2>&1 | tee gpurun_out/r02aj_roialign_alone.log
for t in old new old new old new; do
  if [ $t = old ]; then export MYOLO_LIB=$OLD; else unset MYOLO_LIB; fi
  timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse > gpurun_out/r02aj_bench_$t.json 2> gpurun_out/r02aj_bench_$t.err
  echo "lib=$t $(python - <<PY
import json
d=json.loads(open('gpurun_out/r02aj_bench_$t.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'])
PY
)" | tee -a gpurun_out/r02aj_ab.log
done
unset MYOLO_LIB
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x 2>&1 | tail -3

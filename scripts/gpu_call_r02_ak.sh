#!/bin/bash
# yolo-branch backward started behind the yolo loss (next to the mask head's forward AND backward): tests + A/B (MYOLO_Y_EARLY=0/1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py tests/test_api_gpu.py -q -m gpu -x 2>&1 | tail -3
for f in 0 1 0 1 0 1; do
  MYOLO_Y_EARLY=$f timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02ak_bench_$f.json 2> gpurun_out/r02ak_bench_$f.err
  echo "y_early=$f $(python - <<PY
import json
d=json.loads(open('gpurun_out/r02ak_bench_$f.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], 'sparse', d.get('sparse_backward',{}).get('value'))
PY
)" | tee -a gpurun_out/r02ak_ab.log
done

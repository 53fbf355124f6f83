#!/bin/bash
# conv1 statistics epilogue with the pivot applied after the column reduction: tests + kernel-level timing in the step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_h16_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x 2>&1 | tail -3
for i in 1 2 3; do
  timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run $i', round(d['value'],1), d['ms_per_step'])"
done | tee gpurun_out/r02ae_bench.log

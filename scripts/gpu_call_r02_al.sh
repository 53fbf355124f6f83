#!/bin/bash
# conv1 filter gradient rewritten (lane = pixel, 27 x 4 accumulators): test, kernel time in one step (ncu), step timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv1" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv1_wgrad -c 3 --csv --log-file gpurun_out/r02al_conv1_wgrad.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-fp32-class --no-sparse > /dev/null 2>&1
grep conv1_wgrad gpurun_out/r02al_conv1_wgrad.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-160
for i in 1 2 3; do
  timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run $i', round(d['value'],1), d['ms_per_step'])"
done | tee gpurun_out/r02al_bench.log

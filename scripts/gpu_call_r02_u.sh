#!/bin/bash
# half bn1 path (MYOLO_FUSE_BN bit 16): kernel tests, engine tests, A/B bench 13 vs 29, launch list of the new default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_h16_gpu.py -q -m gpu -x 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py tests/test_api_gpu.py -q -m gpu -x -s 2>&1 | grep -E "passed|failed|Error|error|h16\]" | tail -25
for f in 13 29 13 29; do
  MYOLO_FUSE_BN=$f python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02u_bench_$f.json 2> gpurun_out/r02u_bench_$f.err
  echo "fuse=$f $(python - <<PY
import json
d=json.loads(open('gpurun_out/r02u_bench_$f.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], 'sparse', d.get('sparse_backward',{}).get('value'), 'launches', d.get('gpu_launches'))
PY
)" | tee -a gpurun_out/r02u_ab.log
done

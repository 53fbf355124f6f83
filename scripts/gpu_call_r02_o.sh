#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -k "sparse or replayed" 2>&1 | tail -12 | tee gpurun_out/r02o_sparse_tests.log
python bench.py --no-cpu-baseline --no-parity --no-fp32-class > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; tail -3 gpurun_out/r02o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1])
print('dense', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print('sparse', json.dumps(d['sparse_backward'])[-700:])
PY

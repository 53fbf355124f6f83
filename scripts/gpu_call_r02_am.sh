#!/bin/bash
# programmatic dependent launch for the backbone's kernel chain (MYOLO_PDL=1): full GPU suite with it on, step A/B
mkdir -p gpurun_out
MYOLO_PDL=1 timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02am_tests_pdl.log
for f in 0 1 0 1 0 1; do
  MYOLO_PDL=$f timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02am_bench_$f.json 2> gpurun_out/r02am_bench_$f.err
  echo "pdl=$f $(python - <<PY
import json
d=json.loads(open('gpurun_out/r02am_bench_$f.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], 'sparse', d.get('sparse_backward',{}).get('value'))
PY
)" | tee -a gpurun_out/r02am_ab.log
done

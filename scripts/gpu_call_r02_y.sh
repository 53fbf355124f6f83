#!/bin/bash
# half d(x0) into ROIAlign's backward: tests, sanitizer over the new kernels, A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "roialign" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py tests/test_api_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/sanitize_memcheck_r02y.log \
    python -m pytest tests/test_h16_gpu.py tests/test_kernels_gpu.py -q -m gpu -x -k "bn1 or dgrad_with_fused or roialign" 2>&1 | tail -3
tail -2 gpurun_out/sanitize_memcheck_r02y.log
for f in 0 1 0 1; do
  MYOLO_DX0_HALF=$f timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse > gpurun_out/r02y_bench_$f.json 2> gpurun_out/r02y_bench_$f.err
  echo "dx0_half=$f $(python - <<PY
import json
d=json.loads(open('gpurun_out/r02y_bench_$f.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'])
PY
)" | tee -a gpurun_out/r02y_ab.log
done

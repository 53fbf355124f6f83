#!/bin/bash
# fused BN-backward epilogue with the activation by TMA (MYOLO_WIN_BO=512: per-lane loads as before)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_h16_gpu.py -q -m gpu -x -k "bn or dgrad" 2>&1 | tail -3
for bo in 512 0 512 0; do
  MYOLO_WIN_BO=$bo timeout 120 python scripts/profile_conv.py 4704 20 dgrad_h | tail -1 | sed "s/^/win_bo=$bo /"
done 2>&1 | tee gpurun_out/r02x_dgrad_alone.log
timeout 900 python -m pytest tests/test_h16_gpu.py tests/test_model_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x 2>&1 | tail -3
for f in 512 0 512 0; do
  MYOLO_WIN_BO=$f timeout 600 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02x_bench_$f.json 2> gpurun_out/r02x_bench_$f.err
  echo "win_bo=$f $(python - <<PY
import json
d=json.loads(open('gpurun_out/r02x_bench_$f.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], 'sparse', d.get('sparse_backward',{}).get('value'))
PY
)" | tee -a gpurun_out/r02x_ab.log
done

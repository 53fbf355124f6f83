#!/bin/bash
# A/B of the backbone BN fusions (MYOLO_FUSE_BN bit mask), conv_23 on tcgen05, whole suite with and without fusion
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02d_tests.log
MYOLO_FUSE_BN=7 timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py tests/test_api_gpu.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02d_tests_fused.log
for m in 0 1 4 5 2 7; do
  MYOLO_FUSE_BN=$m python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02d_bench_fuse$m.json 2> gpurun_out/r02d_bench_fuse$m.err
  echo "fuse=$m $(cut -c1-140 gpurun_out/r02d_bench_fuse$m.json)" | tee -a gpurun_out/r02d_ab.log
done
bash scripts/profile_step.sh h16 > gpurun_out/r02d_profile.log 2>&1; head -24 gpurun_out/step_breakdown_h16.txt

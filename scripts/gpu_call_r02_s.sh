#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm or conv3x3 or k_segments" 2>&1 | tail -4 | tee gpurun_out/r02s_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x 2>&1 | tail -4 | tee -a gpurun_out/r02s_tests.log
for t in 1 0 1 0; do
  MYOLO_GEMM_ROW_STORES=$t python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse > gpurun_out/r02s_bench_$t.json 2> gpurun_out/r02s_bench_$t.err
  echo "row_stores=$t $(cut -c1-140 gpurun_out/r02s_bench_$t.json)" | tee -a gpurun_out/r02s_ab.log
done
tail -2 gpurun_out/r02s_bench_0.err
bash scripts/profile_step.sh h16 > gpurun_out/r02s_profile.log 2>&1; grep -E "tc_gemm|launches" gpurun_out/step_breakdown_h16.txt

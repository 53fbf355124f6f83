#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "k_segments or gemm_epilogue" 2>&1 | tail -8 | tee gpurun_out/r02q_tests.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x 2>&1 | tail -4 | tee -a gpurun_out/r02q_tests.log
for t in 0 1 0 1; do
  MYOLO_PW_WIN=$t python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse > gpurun_out/r02q_bench_w$t.json 2> gpurun_out/r02q_bench_w$t.err
  echo "pw_win=$t $(cut -c1-140 gpurun_out/r02q_bench_w$t.json)" | tee -a gpurun_out/r02q_ab.log
done
tail -2 gpurun_out/r02q_bench_w1.err

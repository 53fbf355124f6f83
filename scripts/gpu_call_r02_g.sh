#!/bin/bash
# YOLO-branch backward on its own stream next to the mask-head backward: correctness + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py tests/test_api_gpu.py tests/test_h16_gpu.py -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/r02g_tests.log
for y in 0 1 0 1; do
  MYOLO_Y_OVERLAP=$y python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02g_bench_y$y.json 2> gpurun_out/r02g_bench_y$y.err
  echo "y_overlap=$y $(cut -c1-140 gpurun_out/r02g_bench_y$y.json)" | tee -a gpurun_out/r02g_ab.log
done
tail -3 gpurun_out/r02g_bench_y1.err

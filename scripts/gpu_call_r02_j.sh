#!/bin/bash
# mask-head filter gradients on their own stream behind the data-gradient chain: correctness + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py tests/test_api_gpu.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/r02j_tests.log
for w in 0 1 0 1; do
  MYOLO_W_OVERLAP=$w python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02j_bench_w$w.json 2> gpurun_out/r02j_bench_w$w.err
  echo "w_overlap=$w $(cut -c1-140 gpurun_out/r02j_bench_w$w.json)" | tee -a gpurun_out/r02j_ab.log
done
tail -3 gpurun_out/r02j_bench_w1.err

#!/bin/bash
# strip depthwise kernels: correctness, isolated HBM rates (new vs MYOLO_DW_TILE=1), step A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dwconv" 2>&1 | tail -6 | tee gpurun_out/r02k_dw_tests.log
python scripts/bench_hbm_kernels.py > gpurun_out/r02k_hbm_strip.log 2>&1; grep -E "^dw_.*(dw1 |dw2 |dw3 |dw5 |dw8 |dw14 )|total" gpurun_out/r02k_hbm_strip.log | grep -E "^dw_"
MYOLO_DW_TILE=1 python scripts/bench_hbm_kernels.py > gpurun_out/r02k_hbm_tile.log 2>&1; grep -E "total" gpurun_out/r02k_hbm_tile.log | grep -E "^dw_"
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02k_tests.log
for t in 1 0 1 0; do
  MYOLO_DW_TILE=$t python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02k_bench_t$t.json 2> gpurun_out/r02k_bench_t$t.err
  echo "dw_tile=$t $(cut -c1-140 gpurun_out/r02k_bench_t$t.json)" | tee -a gpurun_out/r02k_ab.log
done

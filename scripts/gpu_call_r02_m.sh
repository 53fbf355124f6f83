#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dwconv" 2>&1 | tail -4 | tee gpurun_out/r02m_dw_tests.log
python scripts/bench_hbm_kernels.py > gpurun_out/r02m_hbm.log 2>&1; grep -E "^dw_.*(dw1 |dw2 |dw5 |total)" gpurun_out/r02m_hbm.log
for t in 1 0 1 0; do
  MYOLO_DW_TILE=$t python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02m_bench_t$t.json 2> gpurun_out/r02m_bench_t$t.err
  echo "dw_tile=$t $(cut -c1-140 gpurun_out/r02m_bench_t$t.json)" | tee -a gpurun_out/r02m_ab.log
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_h16_gpu.py -q -m gpu -k "roialign" 2>&1 | tail -3
python scripts/bench_hbm_kernels.py 2>&1 | grep -E "^roialign" 
MYOLO_LIB=$PWD/mask-yolo_b200/lib/alt_old.so python scripts/bench_hbm_kernels.py 2>&1 | grep -E "^roialign"
for t in old new old new; do
  if [ $t = old ]; then export MYOLO_LIB=$PWD/mask-yolo_b200/lib/alt_old.so; else unset MYOLO_LIB; fi
  python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse > gpurun_out/r02t_bench_$t.json 2> gpurun_out/r02t_bench_$t.err
  echo "roialign=$t $(cut -c1-140 gpurun_out/r02t_bench_$t.json)" | tee -a gpurun_out/r02t_ab.log
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_h16_gpu.py -q -m gpu -k "bn or batch_norm or colsum or stats" 2>&1 | tail -4 | tee gpurun_out/r02n_bn_tests.log
python scripts/bench_hbm_kernels.py 2>&1 | grep -E "^bn_.*total" | tee gpurun_out/r02n_hbm_new.log
MYOLO_LIB=$PWD/mask-yolo_b200/lib/alt_bnold.so python scripts/bench_hbm_kernels.py 2>&1 | grep -E "^bn_.*total" | tee gpurun_out/r02n_hbm_old.log
for t in old new old new; do
  if [ $t = old ]; then export MYOLO_LIB=$PWD/mask-yolo_b200/lib/alt_bnold.so; else unset MYOLO_LIB; fi
  python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02n_bench_$t.json 2> gpurun_out/r02n_bench_$t.err
  echo "bn=$t $(cut -c1-140 gpurun_out/r02n_bench_$t.json)" | tee -a gpurun_out/r02n_ab.log
done
unset MYOLO_LIB
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_config_parity_gpu.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02n_tests.log

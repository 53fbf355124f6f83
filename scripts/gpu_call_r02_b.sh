#!/bin/bash
# tensor-core mask tail: kernel tests, config parity (c5 takes the fused path now), deconv timing A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_h16_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "deconv_mask" -x 2>&1 | tail -15 | tee gpurun_out/r02b_tail_tests.log
timeout 900 python -m pytest tests/test_config_parity_gpu.py tests/test_model_gpu.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02b_parity.log; tail -25 gpurun_out/r02b_parity.log
python scripts/profile_conv.py 4704 20 deconv_h | tail -1 | tee gpurun_out/r02b_deconv.log
MYOLO_MASK_TAIL=ffma python scripts/profile_conv.py 4704 20 deconv_h | tail -1 | tee -a gpurun_out/r02b_deconv.log
python bench.py --no-cpu-baseline --no-parity --no-fp32-class > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; cut -c1-400 gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err

#!/bin/bash
# fused backbone BN (dw BN-on-load + epilogue statistics, GEMM epilogue statistics) + reordered tensor-core mask tail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "dwconv or gemm_epilogue or deconv_mask" 2>&1 | tail -15 | tee gpurun_out/r02c_kernel_tests.log
python scripts/profile_conv.py 4704 20 deconv_h | tail -1 | tee gpurun_out/r02c_deconv.log
MYOLO_MASK_TAIL=ffma python scripts/profile_conv.py 4704 20 deconv_h | tail -1 | tee -a gpurun_out/r02c_deconv.log
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r02c_tests.log
python bench.py --no-cpu-baseline --no-parity --no-fp32-class > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; cut -c1-300 gpurun_out/r02c_bench.json; tail -3 gpurun_out/r02c_bench.err
MYOLO_FUSE_BN=0 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02c_bench_nofuse.json 2> gpurun_out/r02c_bench_nofuse.err; cut -c1-300 gpurun_out/r02c_bench_nofuse.json
bash scripts/profile_step.sh h16 > gpurun_out/r02c_profile.log 2>&1; head -30 gpurun_out/step_breakdown_h16.txt

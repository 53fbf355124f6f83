#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_h16_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "deconv_mask" 2>&1 | tail -3
python scripts/profile_conv.py 4704 20 deconv_h | tail -1
python scripts/profile_conv.py 4704 20 deconv | tail -1

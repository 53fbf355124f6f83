#!/bin/bash
# compute-sanitizer memcheck over the kernels touched at the end of round 2 (conv1 filter gradient, ROIAlign forward cache, bn1 half passes,
# mask tail, fused BN backward with the TMA activation, PDL launches) + one sanitized training step at a small size
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/sanitize_memcheck_final.log \
    python -m pytest tests/test_kernels_gpu.py tests/test_h16_gpu.py -q -m gpu -x -k "conv1 or roialign or bn1 or dgrad_with_fused or deconv or dw or bn_backward or wgrad" 2>&1 | tail -3
tail -2 gpurun_out/sanitize_memcheck_final.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/sanitize_memcheck_step.log \
    python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "replayed or sparse" 2>&1 | tail -3
tail -2 gpurun_out/sanitize_memcheck_step.log

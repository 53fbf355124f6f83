#!/bin/bash
# compute-sanitizer over the per-kernel GPU tests (race detection / sanitizers).  Slow (10-50x): run a subset.
# usage (under gpurun, one GPU):  bash scripts/sanitize.sh [memcheck|racecheck|initcheck|synccheck] [pytest -k expression]
TOOL=${1:-memcheck}
EXPR=${2:-"dw or roialign or decode or targets or loss or shapes or postprocess"}
mkdir -p gpurun_out
compute-sanitizer --tool ${TOOL} --error-exitcode 3 --log-file gpurun_out/sanitize_${TOOL}.log \
    python -m pytest tests/test_kernels_gpu.py tests/test_shapes_raster.py -q -m gpu -x -k "${EXPR}" 2>&1 | tail -5
echo "exit code: $?"; grep -c "ERROR SUMMARY" gpurun_out/sanitize_${TOOL}.log; tail -3 gpurun_out/sanitize_${TOOL}.log

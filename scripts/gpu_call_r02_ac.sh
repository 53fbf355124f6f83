#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python -m pytest "tests/test_api_gpu.py::test_train_loop_checkpoint_and_reload" -q -m gpu -x 2>&1 | grep -E "assert|Error|passed|failed|hist|\[" | head -12
done
timeout 1500 python -m pytest tests/ -q -m gpu 2>&1 | tail -6

#!/bin/bash
# full GPU suite on the current build + per-kernel roofline table of one step + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -4
bash scripts/profile_step_metrics.sh h16 c2 > /dev/null 2>&1
head -45 gpurun_out/step_metrics_h16_c2.txt

#!/bin/bash
# ncu launch list of one training step (all kernels, serialised) -> gpurun_out/launches_step.csv
# usage (under gpurun): bash scripts/profile_step.sh [precision]
PREC=${1:-h16}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_${PREC}.csv \
    python bench.py --precision ${PREC} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-fp32-class --no-sparse > gpurun_out/launches_step_${PREC}.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_step_${PREC}.csv > gpurun_out/step_breakdown_${PREC}.txt
tail -60 gpurun_out/step_breakdown_${PREC}.txt

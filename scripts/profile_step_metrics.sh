#!/bin/bash
# One training step at the benchmark configuration under ncu with the roofline metrics of EVERY launch: duration, DRAM
# bytes read / written, DRAM throughput %, tensor-pipe % -> gpurun_out/step_metrics_<prec>.csv and a per-kernel table.
# (cold-cache, serialised: compare shares and per-kernel rates, not the step total)
# usage (under gpurun, one GPU): bash scripts/profile_step_metrics.sh [precision] [config]
PREC=${1:-h16}
CFG=${2:-c2}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/step_metrics_${PREC}_${CFG}.csv \
    python bench.py --config ${CFG} --precision ${PREC} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-fp32-class --no-sparse \
    > gpurun_out/step_metrics_${PREC}_${CFG}.log 2>&1
python scripts/summarize_step_metrics.py gpurun_out/step_metrics_${PREC}_${CFG}.csv > gpurun_out/step_metrics_${PREC}_${CFG}.txt
head -70 gpurun_out/step_metrics_${PREC}_${CFG}.txt

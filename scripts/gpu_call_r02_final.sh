#!/bin/bash
# final state of round 2: whole GPU suite, smoke, the three benchmark lines, the reference arm header, per-kernel metrics of one step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/final_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
python bench.py > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err; cut -c1-300 gpurun_out/bench_r02_c2.json; tail -2 gpurun_out/bench_r02_c2.err
python bench.py --config c3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; cut -c1-200 gpurun_out/bench_r02_c3.json; tail -2 gpurun_out/bench_r02_c3.err
python bench.py --config c5 > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; cut -c1-200 gpurun_out/bench_r02_c5.json; tail -2 gpurun_out/bench_r02_c5.err
bash scripts/profile_step_metrics.sh h16 c2 > gpurun_out/final_metrics.log 2>&1; head -30 gpurun_out/step_metrics_h16_c2.txt

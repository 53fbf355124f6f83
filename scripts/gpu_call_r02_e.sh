#!/bin/bash
# suite, c3 / c5 bench lines, per-kernel roofline metrics of one step, ncu --set full of the dominant kernel -> traffic json
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02e_tests.log
python bench.py --config c3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; cut -c1-500 gpurun_out/bench_r02_c3.json; tail -3 gpurun_out/bench_r02_c3.err
python bench.py --config c5 > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; cut -c1-500 gpurun_out/bench_r02_c5.json; tail -3 gpurun_out/bench_r02_c5.err
bash scripts/profile_step_metrics.sh h16 c2 > gpurun_out/r02e_metrics.log 2>&1; head -45 gpurun_out/step_metrics_h16_c2.txt
bash scripts/capture_ncu.sh fwd_h fwd
python scripts/make_roofline_traffic.py gpurun_out/ncu_fwd_h.raw.csv 4704 gpurun_out/ncu_fwd.raw.csv
cp profiles/roofline_traffic.json gpurun_out/roofline_traffic.json
rm -f gpurun_out/*.ncu-rep

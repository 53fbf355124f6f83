#!/bin/bash
# final state of round 2 (third pass): whole GPU suite (stop if it fails), smoke, conv1 filter-gradient kernel time, the three benchmark lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final3_tests.log
grep -q "failed\|error" gpurun_out/final3_tests.log && { echo "GPU suite not green: stopping"; exit 1; }
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final3_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv1_wgrad -c 3 --csv --log-file gpurun_out/final3_conv1_wgrad.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-fp32-class --no-sparse > /dev/null 2>&1
grep conv1_wgrad gpurun_out/final3_conv1_wgrad.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-140
python bench.py > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err; cut -c1-300 gpurun_out/bench_r02_c2.json; tail -2 gpurun_out/bench_r02_c2.err
python bench.py --config c3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; cut -c1-200 gpurun_out/bench_r02_c3.json; tail -2 gpurun_out/bench_r02_c3.err
python bench.py --config c5 > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; cut -c1-200 gpurun_out/bench_r02_c5.json; tail -2 gpurun_out/bench_r02_c5.err

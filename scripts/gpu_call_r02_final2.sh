#!/bin/bash
# final state of round 2 (second pass): whole GPU suite, smoke, the three benchmark lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final2_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final2_smoke.log
python bench.py > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err; cut -c1-300 gpurun_out/bench_r02_c2.json; tail -2 gpurun_out/bench_r02_c2.err
python bench.py --config c3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; cut -c1-200 gpurun_out/bench_r02_c3.json; tail -2 gpurun_out/bench_r02_c3.err
python bench.py --config c5 > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; cut -c1-200 gpurun_out/bench_r02_c5.json; tail -2 gpurun_out/bench_r02_c5.err

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02i_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02i_smoke.log
python bench.py > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err; cut -c1-400 gpurun_out/bench_r02_c2.json; tail -3 gpurun_out/bench_r02_c2.err

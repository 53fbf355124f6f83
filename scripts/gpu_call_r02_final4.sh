#!/bin/bash
# final state of round 2 (fourth pass, after the stream-schedule change and the polygon rasteriser): whole GPU suite (stop if it
# fails), smoke, polygon rasteriser timing, the three benchmark lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final4_tests.log
grep -q "failed\|error" gpurun_out/final4_tests.log && { echo "GPU suite not green: stopping"; exit 1; }
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/final4_smoke.log
python scripts/polygon_raster_time.py 2>&1 | tail -1 | cut -c1-400
python bench.py > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err; cut -c1-300 gpurun_out/bench_r02_c2.json; tail -2 gpurun_out/bench_r02_c2.err
python bench.py --config c3 > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; cut -c1-200 gpurun_out/bench_r02_c3.json; tail -2 gpurun_out/bench_r02_c3.err
python bench.py --config c5 > gpurun_out/bench_r02_c5.json 2> gpurun_out/bench_r02_c5.err; cut -c1-200 gpurun_out/bench_r02_c5.json; tail -2 gpurun_out/bench_r02_c5.err

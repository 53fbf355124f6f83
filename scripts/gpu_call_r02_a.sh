#!/bin/bash
# round-2 first GPU call: whole GPU suite, smoke, default bench line (both precisions, e2e, parity at the headline config),
# reference arm at the driver's K/W, launch list of one h16 step, HBM-kernel micro-benchmarks
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rxXs --durations=15 -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r02a_tests.log; tail -40 gpurun_out/r02a_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02a_smoke.log
python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
bash scripts/profile_step.sh h16 > gpurun_out/r02a_profile.log 2>&1; tail -5 gpurun_out/r02a_profile.log
python scripts/bench_hbm_kernels.py > gpurun_out/r02a_hbm.log 2>&1; tail -12 gpurun_out/r02a_hbm.log
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err; tail -c 600 gpurun_out/r02a_ref.json; tail -4 gpurun_out/r02a_ref.err

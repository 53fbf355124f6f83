#!/bin/bash
# One gpurun call that re-establishes the measured state of the repository on a fresh B200 (about 3 minutes of box time):
#   1. the whole GPU test suite (includes the xfail-non-strict tests that round 1 could not run: watch for XPASS/XFAIL)
#   2. smoke() and the default bench line (device-timed value, e2e, roofline, cpu_baseline, parity)
#   3. the ncu launch list of one step and its per-kernel breakdown
#   4. timing of the device Shapes generator and the device-fed end-to-end rate
# usage:  gpurun --timeout 420 -- 'bash scripts/first_gpu_call.sh'
# with two GPUs (gpurun --gpus 2) it also runs the two-rank DDP tests, including the C-ABI NCCL transport.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rxX 2>&1 | tail -15 | tee gpurun_out/first_tests.log
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/first_smoke.log
python bench.py > gpurun_out/first_bench.json 2> gpurun_out/first_bench.err; tail -c 1500 gpurun_out/first_bench.json
bash scripts/profile_step.sh h16 | tail -25
python scripts/shapes_raster_time.py | tail -1
python scripts/device_feed_e2e.py | tail -1
python scripts/bench_hbm_kernels.py | tail -12
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --no-cpu-baseline > gpurun_out/first_bench_2gpu.json 2> gpurun_out/first_bench_2gpu.err
  tail -c 600 gpurun_out/first_bench_2gpu.json
fi

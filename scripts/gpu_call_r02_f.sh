#!/bin/bash
# two GPUs: DDP tests (both transports), 2-rank bench with NCCL debug summary; bn1 epilogue statistics A/B on one GPU
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_ddp_gpu.py tests/test_h16_gpu.py -q -m gpu -k "ddp or two_gpu or cabi or conv3x3_half" -rs 2>&1 | tail -8 | tee gpurun_out/r02f_ddp_tests.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --no-cpu-baseline --no-parity > gpurun_out/r02f_bench_2gpu.out 2> gpurun_out/r02f_bench_2gpu.err
grep '"metric"' gpurun_out/r02f_bench_2gpu.out > gpurun_out/bench_r02_2gpu.json; cut -c1-600 gpurun_out/bench_r02_2gpu.json
grep -E "NCCL INFO (Channel|Connected|comm|Using|NVLS|ncclCommInit)|via P2P|NET/" gpurun_out/r02f_bench_2gpu.out gpurun_out/r02f_bench_2gpu.err | head -30 > gpurun_out/r02f_nccl_info.txt; wc -l gpurun_out/r02f_nccl_info.txt
for m in 5 13; do
  MYOLO_FUSE_BN=$m python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e > gpurun_out/r02f_bench_fuse$m.json 2> gpurun_out/r02f_bench_fuse$m.err
  echo "fuse=$m $(cut -c1-140 gpurun_out/r02f_bench_fuse$m.json)" | tee -a gpurun_out/r02f_ab.log
done

#!/bin/bash
# eight GPUs: two-rank DDP tests with the stream-overlapped backward, then the weak-scaling bench line at N = 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_ddp_gpu.py -q -m gpu -rs 2>&1 | tail -4 | tee gpurun_out/r02p_ddp_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --no-cpu-baseline --no-parity > gpurun_out/r02p_bench_8gpu.out 2> gpurun_out/r02p_bench_8gpu.err
grep '"metric"' gpurun_out/r02p_bench_8gpu.out > gpurun_out/bench_r02_8gpu.json; tail -3 gpurun_out/r02p_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_8gpu.json').read().strip().splitlines()[-1])
print('N=8 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'sync', d['e2e_train_on_batch']['value'])
print('fp32', d['fp32_class']['value'], d['fp32_class']['e2e']['value'], 'sparse', d['sparse_backward']['value'], d['sparse_backward']['e2e']['value'])
PY

#!/bin/bash
# two GPUs on the final stream schedule: the two-rank DDP tests (both transports), then the weak-scaling bench line at N = 2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_ddp_gpu.py -q -m gpu -rs 2>&1 | tail -4 | tee gpurun_out/r02aw_ddp_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 2 --no-cpu-baseline --no-parity --no-fp32-class --no-sparse > gpurun_out/r02aw_bench_2gpu.out 2> gpurun_out/r02aw_bench_2gpu.err
grep '"metric"' gpurun_out/r02aw_bench_2gpu.out > gpurun_out/bench_r02_2gpu.json; tail -3 gpurun_out/r02aw_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_2gpu.json').read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'sync', d['e2e_train_on_batch']['value'])
PY

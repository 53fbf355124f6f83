#!/bin/bash
# fused BN-backward epilogue: early activation fetch (MYOLO_WIN_BO=64 switches it off) -- tests, standalone timing, ncu source capture, bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_h16_gpu.py -q -m gpu -x -k "bn or dgrad" 2>&1 | tail -3
for i in 1 2; do
  MYOLO_WIN_BO=64 python scripts/profile_conv.py 4704 20 dgrad_h | tail -1 | sed 's/^/early_fetch=off /'
  python scripts/profile_conv.py 4704 20 dgrad_h | tail -1 | sed 's/^/early_fetch=on  /'
done 2>&1 | tee gpurun_out/r02v_dgrad_alone.log
python scripts/profile_conv.py 4704 20 fwd_h | tail -1 | tee -a gpurun_out/r02v_dgrad_alone.log
ncu --set full --clock-control none --import-source on -k regex:'tc_conv_win_kernel' -s 2 -c 1 -f -o gpurun_out/ncu_dgrad_h_r02v python scripts/profile_conv.py 4704 4 dgrad_h > gpurun_out/ncu_dgrad_h_r02v.log 2>&1
ncu -i gpurun_out/ncu_dgrad_h_r02v.ncu-rep --page raw --csv > gpurun_out/ncu_dgrad_h_r02v.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_dgrad_h_r02v.ncu-rep --page source --csv > gpurun_out/ncu_dgrad_h_r02v.source.csv 2>/dev/null
for f in 64 0 64 0; do
  MYOLO_WIN_BO=$f python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse > gpurun_out/r02v_bench_$f.json 2> gpurun_out/r02v_bench_$f.err
  echo "win_bo=$f $(python - <<PY
import json
d=json.loads(open('gpurun_out/r02v_bench_$f.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'])
PY
)" | tee -a gpurun_out/r02v_ab.log
done

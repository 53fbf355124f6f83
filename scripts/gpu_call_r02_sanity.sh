#!/bin/bash
# convergence sanity of the final build: 400 steps on Shapes 224 from the Keras default init, dense and sparse-backward h16, tf32x3
mkdir -p gpurun_out
for mode in "h16 0" "h16 1" "tf32x3 0"; do
  set -- $mode
  echo "== precision $1  MYOLO_SPARSE_BWD=$2"
  MYOLO_SPARSE_BWD=$2 python scripts/train_sanity.py 224 400 $1 2>&1 | awk 'NR%2==1 || /steps in/'
done | tee gpurun_out/r02_train_sanity.txt

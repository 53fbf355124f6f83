#!/bin/bash
# phase timeline of the step: default | filter gradients deferred to their own stream on all / 110 / 92 SMs
mkdir -p gpurun_out
bash scripts/phase_timeline.sh "MYOLO_NOP=1" "MYOLO_W_OVERLAP=1" "MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SMS=110" "MYOLO_W_OVERLAP=1 MYOLO_WGRAD_SMS=92" 2>&1 | tee gpurun_out/r02as_phases.log

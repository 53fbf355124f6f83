#!/bin/bash
# What does each kernel family cost ON THE CRITICAL PATH of the training step?  The step is timed with the named C-ABI entry
# points switched off (MYOLO_WHATIF_SKIP, results garbage, timing valid for the data-independent kernels that remain) and
# compared with the full step.  Usage (under gpurun): bash scripts/whatif_critical_path.sh > gpurun_out/whatif.txt
mkdir -p gpurun_out
run() {
  MYOLO_WHATIF_SKIP="$1" timeout 300 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-70s %8.3f ms/step  %d launches' % ('$1' or '(full step)', d['ms_per_step'], d['gpu_launches']//d['steps']))"
}
run ""
run "myolo_conv1_wgrad"
run "myolo_roialign_bwd_h"
run "myolo_roialign_fwd_h"
run "myolo_bn_apply_hh"
run "myolo_bn_bwd_batch_fix_hh"
run "myolo_mask_out_bwd_h"
run "myolo_bn_bwd"
run "myolo_bn_apply,myolo_bn_apply_split"
run "myolo_dwconv3x3_bwd_filter,myolo_dwconv3x3_bwd_filter_bn"
run "myolo_dwconv3x3_bwd_data"
run "myolo_dwconv3x3_fwd_bn,myolo_dwconv3x3_fwd"
run "myolo_pwconv_wgrad"
run "myolo_pwconv_dgrad"
run "myolo_gemm_taps_tc_stats,myolo_gemm_taps_tc"
run "myolo_gemm_taps_wgrad_h"
run "myolo_conv3x3_wgrad,myolo_conv3x3_dgrad"
run "myolo_adam_step,myolo_adam_step_masked"
run ""

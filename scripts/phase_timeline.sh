#!/bin/bash
# Main-stream phase boundaries of one training step (MYOLO_PHASES=1: timing events recorded inside the step), for a list of
# environment settings.  Usage: scripts/phase_timeline.sh "ENV1=a ENV2=b" "ENV3=c" ...   (run on the GPU box)
for envs in "$@"; do
  echo "== $envs"
  env MYOLO_PHASES=1 $envs timeout 300 python bench.py --no-cpu-baseline --no-parity --no-fp32-class --no-e2e --no-sparse 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  step %.3f ms' % d['ms_per_step'])
for n,ms in d['phases_ms']: print('  %7.3f ms  %s' % (ms, n))
print('  %7.3f ms  (sum)' % sum(ms for _,ms in d['phases_ms']))"
done

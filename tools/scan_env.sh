#!/bin/bash
# Usage: tools/scan_env.sh VAR "v1 v2 ..." [extra env assignments]   -- bench cfg4 for each value of VAR
VAR=$1; VALS=$2; shift 2
for v in $VALS; do
  env $VAR=$v "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$VAR=$v $*', 'ms/step %.3f' % d['ms_per_step'], {k: round(x,2) for k,x in d['roofline']['kernel_ms_per_step'].items()})"
done

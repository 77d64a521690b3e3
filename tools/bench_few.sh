#!/bin/bash
# tools/bench_few.sh "cfgA cfgB ..." [bench flags]
CFGS=$1; shift
for c in $CFGS; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-checksum "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-14s %.3e %s  ms/step %.3f' % ('$c', d['value'], d['unit'], d['ms_per_step']), {k: round(x,2) for k,x in d['roofline']['kernel_ms_per_step'].items()})"
done

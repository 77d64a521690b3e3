#!/bin/bash
# Usage: tools/ncu_launches.sh <out-name>   -- plain run, then the ncu launch list of the same command
set -u
OUT=$1
B="python bench.py --steps 2 --warmup 1 --elems 67108864 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_${OUT}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${OUT}.csv $B > gpurun_out/ncu_${OUT}.log 2>&1
tail -1 gpurun_out/ncu_${OUT}.log | cut -c1-200

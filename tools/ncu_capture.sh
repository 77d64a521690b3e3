#!/bin/bash
# Usage: tools/ncu_capture.sh <out-name> <kernel-regex> [skip] [count]
# Plain run first (must exit 0), then one ncu --set full capture of the matching kernels.
set -u
OUT=$1; REGEX=$2; SKIP=${3:-0}; COUNT=${4:-3}
B="python bench.py --steps 1 --warmup 1 --elems 33554432 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_${OUT}.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:${REGEX}" -s ${SKIP} -c ${COUNT} -o gpurun_out/${OUT} $B > gpurun_out/ncu_${OUT}.log 2>&1
tail -2 gpurun_out/ncu_${OUT}.log

#!/bin/bash
# Usage: tools/ncu_case.sh <out-name> <case> <kernel-regex> [skip] [count]
set -u
OUT=$1; CASE=$2; REGEX=$3; SKIP=${4:-0}; COUNT=${5:-2}
B="python tools/profile_case.py $CASE"
$B > gpurun_out/plain_${OUT}.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:${REGEX}" -s ${SKIP} -c ${COUNT} -o gpurun_out/${OUT} $B > gpurun_out/ncu_${OUT}.log 2>&1
tail -1 gpurun_out/ncu_${OUT}.log

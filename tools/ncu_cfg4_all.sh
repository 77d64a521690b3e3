#!/bin/bash
# One cfg4 pass (2^25 elements) under ncu --set full: every kernel of libpgm_cuda.so once.
set -u
OUT=${1:-prof_cfg4_all}
B="python tools/profile_case.py cfg4 33554432 2"
$B > gpurun_out/plain_${OUT}.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:pgm::(bucket|setup|sample|dense)" -s 13 -c 13 -o gpurun_out/${OUT} $B > gpurun_out/ncu_${OUT}.log 2>&1
tail -1 gpurun_out/ncu_${OUT}.log

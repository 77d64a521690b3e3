#!/bin/bash
# A/B of prebuilt library variants on the GPU box: tools/ab_libs.sh "cfgA cfgB" name1 name2 ...   (build_variants/lib_<name>.so)
CFGS=$1; shift
cp polyagamma_b200/libpgm_cuda.so /tmp/lib_current.so
for name in "$@"; do
  cp build_variants/lib_${name}.so polyagamma_b200/libpgm_cuda.so
  echo "=== $name"
  bash tools/bench_few.sh "$CFGS" --steps 4
done
cp /tmp/lib_current.so polyagamma_b200/libpgm_cuda.so

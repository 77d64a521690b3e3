#!/bin/bash
# A/B of prebuilt library variants (build_variants/*.so): tools/perf_configs.py lines matching $1
for f in build_variants/*.so; do
  name=$(basename $f .so)
  cp $f polyagamma_b200/libpgm_cuda.so
  echo "=== $name"
  python tools/perf_configs.py 2>&1 | grep -E "$1" | cut -c1-${2:-60}
done
cp build_variants/lib_default.so polyagamma_b200/libpgm_cuda.so

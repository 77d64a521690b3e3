#!/bin/bash
# A/B of prebuilt library variants (build_variants/*.so): bench.py kernel breakdown only
for f in build_variants/*.so; do
  name=$(basename $f .so)
  cp $f polyagamma_b200/libpgm_cuda.so
  python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$name', round(d['value']/1e9,3), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
done
cp build_variants/lib_default.so polyagamma_b200/libpgm_cuda.so

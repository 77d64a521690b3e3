#!/bin/bash
# A/B of prebuilt library variants (build_variants/*.so) on the GPU box: perf of every config
for f in build_variants/*.so; do
  name=$(basename $f .so)
  cp $f polyagamma_b200/libpgm_cuda.so
  echo "=== $name"
  python tools/perf_configs.py 2>&1 | cut -c1-56
  cp gpurun_out/perf_configs.json gpurun_out/perf_${name}.json
  python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('bench', d['value'], d['ms_per_step'], {k: round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
done
cp build_variants/lib_default.so polyagamma_b200/libpgm_cuda.so

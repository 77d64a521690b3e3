#!/bin/bash
# One bench.py line per BASELINE config into gpurun_out/r2/bench_<config>.json
# Usage: tools/bench_all.sh [extra bench.py flags]
mkdir -p gpurun_out/r2
for c in cfg1 cfg2 cfg2-scalar-h cfg3-devroye cfg3-alternate cfg3-saddle cfg3-gamma cfg4 cfg5a-pdf cfg5a-cdf cfg5a-logpdf cfg5a-logcdf; do
  python bench.py --config $c --steps 5 --warmup 3 "$@" > gpurun_out/r2/bench_$c.json 2> gpurun_out/r2/bench_$c.err
  rc=$?
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2/bench_$c.json"))
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print("%-14s rc=$rc value %.3e %s  ms/step %.3f  e2e %s  dom %s frac %s launches %s" % ("$c", d["value"], d["unit"], d["ms_per_step"],
          ("%.3e" % e["value"]) if e else None, (r.get("kernel") or "")[:28], ("%.3f" % r["frac"]) if r.get("frac") else None, d["gpu_launches"]))
except Exception as ex:
    print("$c rc=$rc FAILED", ex)
    print(open("gpurun_out/r2/bench_$c.err").read()[-600:])
PY
done

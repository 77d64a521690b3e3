#!/bin/bash
# the reference's truncation-point table against the acceptance-optimal one (the same table x 0.5),
# two prebuilt libraries (build_variants/lib_trunc_{ref,half}.so, built by
#   make -C polyagamma_b200/csrc EXTRA=-DPGM_TRUNC_SCALE=0.5).  For each: validity (exact moments of 1e7 draws at
# the cells where the envelope matters most) and speed (explicit alternate at z = 0, the cfg3 alternate sweep).
cp polyagamma_b200/libpgm_cuda.so /tmp/lib_cur.so
for v in ref half; do
  cp build_variants/lib_trunc_$v.so polyagamma_b200/libpgm_cuda.so
  echo "=== truncation table: $v"
  python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import polyagamma_b200 as pg
for h, z in ((1.5, 0.0), (2.5, 0.0), (4.0, 0.0), (3.5, 1.0), (2.0, 2.0)):
    n = 10_000_000
    x = pg.random_polyagamma(h, z, size=n, method="alternate", random_state=(5, 1), device="cuda")
    mean = h * np.tanh(z / 2) / (2 * z) if z else h / 4
    var = h / 24 if z == 0 else h * (np.sinh(z) - z) / np.cosh(z / 2) ** 2 / (4 * z ** 3)
    xm = float(x.mean()); v = float(x.var()); m4 = float(((x - xm) ** 4).mean())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(5):
        pg.random_polyagamma(h, z, size=n, method="alternate", random_state=(5, k), device="cuda")
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("h=%.1f z=%.1f  mean z-score %+.2f  var z-score %+.2f   %.2fe9 samples/s" % (h, z, (xm - mean) / np.sqrt(var / n),
          (v - var) / np.sqrt((m4 - var * var) / n), n / dt / 1e9))
PY
  python bench.py --config cfg3-alternate --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-checksum 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cfg3-alternate %.3e samples/s  ms/step %.3f' % (d['value'], d['ms_per_step']), {k: round(x,2) for k,x in d['roofline']['kernel_ms_per_step'].items()})"
done
cp /tmp/lib_cur.so polyagamma_b200/libpgm_cuda.so

"""First-contact GPU check: stream parity of every method with the oracle, density parity,
rough timings.  Writes gpurun_out/first_check.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import oracle as O
import polyagamma_b200 as pg

res = {}
print(torch.cuda.get_device_name(0), flush=True)
rng = np.random.default_rng(1)
n = 50000
cases = {
    "devroye": (rng.integers(1, 5, n).astype(float), rng.uniform(-50, 50, n)),
    "devroye_smallz": (np.ones(n), rng.normal(0, 3, n)),
    "alternate": (rng.uniform(1, 4, n), rng.uniform(-50, 50, n)),
    "alternate_big": (rng.uniform(0.1, 12, n), rng.uniform(-6, 6, n)),
    "saddle": (rng.uniform(10, 100, n), rng.uniform(-50, 50, n)),
    "saddle_small": (rng.uniform(4.1, 12, n), rng.uniform(-5, 5, n)),
    "gamma": (rng.uniform(0.1, 3, 2000), rng.uniform(-50, 50, 2000)),
    None: (rng.uniform(0.1, 200, n), rng.uniform(-50, 50, n)),
}
for name, (h, z) in cases.items():
    method = None if name is None else name.split("_")[0]
    for compat in (False, True):
        t = time.time()
        got = pg.random_polyagamma(torch.from_numpy(h).cuda(), torch.from_numpy(z).cuda(), method=method,
                                   random_state=(7, 3), ref_compat=compat)
        torch.cuda.synchronize(); dt = time.time() - t
        got = got.cpu().numpy()
        want = O.fill2(7, 3, h, z, method, ref_compat=compat)
        rel = np.abs(got - want) / np.abs(want)
        frac = float(np.mean(rel < 1e-9))
        res[f"{name}_compat{int(compat)}"] = dict(frac_agree=frac, max_rel_agreeing=float(rel[rel < 1e-9].max()),
                                                  nan=int(np.isnan(got).sum()), first_call_s=dt)
        print(name, compat, res[f"{name}_compat{int(compat)}"], flush=True)
# scalar fill
for method, h, z in (("devroye", 1, 0), ("alternate", 2.5, 1), ("saddle", 20, 3), (None, 60, 2), ("gamma", 1.5, 1)):
    nn = 20000 if method != "gamma" else 1000
    got = pg.random_polyagamma(h, z, size=nn, method=method, random_state=(9, 0))
    want = O.fill(9, 0, h, z, nn, method)
    rel = np.abs(got - want) / np.abs(want)
    print("fill", method, float(np.mean(rel < 1e-9)), flush=True)
    res[f"fill_{method}"] = float(np.mean(rel < 1e-9))
# density
x = rng.uniform(0.001, 6, 20000); hh = rng.uniform(0.2, 10, 20000); zz = rng.uniform(-12, 12, 20000); zz[::9] = 0
for name, f, log in (("pdf", pg.polyagamma_pdf, False), ("cdf", pg.polyagamma_cdf, False),
                     ("logpdf", pg.polyagamma_pdf, True), ("logcdf", pg.polyagamma_cdf, True)):
    a = f(x, hh, zz, return_log=log)
    b, cond = O.density(name, x, hh, zz, return_cond=True)
    ok = np.isfinite(b) & (cond < 1e3) & (np.abs(b) > 1e-280)
    err = np.abs(a[ok] - b[ok]) / np.abs(b[ok])
    res["density_" + name] = dict(max_rel=float(err.max()), checked=float(ok.mean()))
    print(name, res["density_" + name], flush=True)
# timings (device resident)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 1e3)
    return min(ts)
N = 1 << 26
g = torch.Generator(device="cuda"); g.manual_seed(0)
zN = torch.randn(N, device="cuda", dtype=torch.float64, generator=g) * 3
hU = torch.rand(N, device="cuda", dtype=torch.float64, generator=g)
zU = torch.rand(N, device="cuda", dtype=torch.float64, generator=g) * 100 - 50
one = torch.ones(N, device="cuda", dtype=torch.float64)
out = torch.empty(N, device="cuda", dtype=torch.float64)
timing = {}
timing["cfg1_fill_h1_z0"] = N / timeit(lambda: pg.random_polyagamma(1.0, 0.0, size=N, device="cuda", random_state=(1, 0)))
timing["cfg2_fill2_h1_zN03"] = N / timeit(lambda: pg.random_polyagamma(one, zN, out=out, random_state=(1, 0), disable_checks=True))
timing["cfg3_devroye"] = N / timeit(lambda: pg.random_polyagamma(one, zU, out=out, method="devroye", random_state=(1, 0), disable_checks=True))
timing["cfg3_alternate"] = N / timeit(lambda: pg.random_polyagamma(1 + 3 * hU, zU, out=out, method="alternate", random_state=(1, 0), disable_checks=True))
timing["cfg3_saddle"] = N / timeit(lambda: pg.random_polyagamma(10 + 90 * hU, zU, out=out, method="saddle", random_state=(1, 0), disable_checks=True))
timing["cfg4_hybrid"] = N / timeit(lambda: pg.random_polyagamma(0.1 + 199.9 * hU, zU, out=out, random_state=(1, 0), disable_checks=True))
M = 1 << 22
timing["cfg3_gamma"] = M / timeit(lambda: pg.random_polyagamma(0.1 + 0.9 * hU[:M], zU[:M], out=out[:M], method="gamma", random_state=(1, 0), disable_checks=True), reps=1)
xg = torch.rand(N, device="cuda", dtype=torch.float64, generator=g) * 12.5 + 1e-3
hg = 0.5 + 24.5 * hU; zg = zU / 5
for name, f, log in (("pdf", pg.polyagamma_pdf, False), ("cdf", pg.polyagamma_cdf, False),
                     ("logpdf", pg.polyagamma_pdf, True), ("logcdf", pg.polyagamma_cdf, True)):
    m = N if name != "logcdf" else N // 8
    timing["density_" + name] = m / timeit(lambda: f(xg[:m], hg[:m], zg[:m], return_log=log), reps=2)
for k, v in timing.items():
    print(f"{k:24s} {v/1e6:12.1f} M/s", flush=True)
res["timing_per_s"] = timing
res["launches"] = pg.launch_count()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "first_check.json"), "w"), indent=1)

"""oracle/ref_bench.py -- TEST/BENCH INFRASTRUCTURE ONLY.

Times the reference's own CPU implementation of the hot path on the host cores, through the reference's
Python API (`random_polyagamma(h, z, out=..., disable_checks=True)` -> pgm_random_polyagamma_fill2 /
_fill with numpy PCG64, polyagamma/_polyagamma.pyx:163-193,244-251; `polyagamma_pdf / polyagamma_cdf` ->
dispatch(), :433-472): persistent worker processes, one per core, each working on a bounded sample of the
bench workload; the main process times barrier-to-barrier, so a step's time is the slowest worker's.
Used only by bench.py (`cpu_baseline` and `--impl reference`).  Falls back to the oracle port
(libpgm_oracle.so) when oracle/_ref has not been built, and says so through `kind`.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# workload -> (h, z, method, density kind); distributions of SURVEY.md 8(d) / BASELINE.json configs
SAMPLERS = {
    "cfg1": ("one", "zero", None),
    "cfg2": ("ones", "normal3", None),
    "cfg2-scalar-h": ("one", "normal3", None),
    "cfg3-devroye": ("ones", "u50", "devroye"),
    "cfg3-alternate": ("u1_4", "u50", "alternate"),
    "cfg3-saddle": ("u10_100", "u50", "saddle"),
    "cfg3-gamma": ("u01_1", "u50", "gamma"),
    "cfg4": ("u01_200", "u50", None),
}
DENSITIES = {"cfg5a-pdf": ("pdf", False), "cfg5a-logpdf": ("pdf", True), "cfg5a-cdf": ("cdf", False),
             "cfg5a-logcdf": ("cdf", True)}
# elements (points) per core and step: about a second of reference work each
DEFAULT_SAMPLE = {"cfg1": 8_000_000, "cfg2": 5_000_000, "cfg2-scalar-h": 5_000_000, "cfg3-devroye": 5_000_000,
                  "cfg3-alternate": 3_000_000, "cfg3-saddle": 1_200_000, "cfg3-gamma": 80_000, "cfg4": 2_000_000,
                  "cfg5a-pdf": 3_000_000, "cfg5a-logpdf": 160_000, "cfg5a-cdf": 1_500_000, "cfg5a-logcdf": 45_000}


def draw_operand(rng, name, n):
    if name == "one":
        return 1.0
    if name == "zero":
        return 0.0
    if name == "ones":
        return np.ones(n)
    if name == "normal3":
        return rng.normal(0.0, 3.0, n)
    if name == "u50":
        return rng.uniform(-50.0, 50.0, n)
    lo, hi = {"u1_4": (1.0, 4.0), "u10_100": (10.0, 100.0), "u01_1": (0.1, 1.0), "u01_200": (0.1, 200.0)}[name]
    return rng.uniform(lo, hi, n)


def make_job(workload: str, m: int, seed: int, use_ref: bool):
    """A callable job(step) doing one step of `workload` on m elements with the reference (or the port)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    if workload in SAMPLERS:
        hn, zn, method = SAMPLERS[workload]
        if workload == "cfg1":
            # BASELINE configs[0]: random_polyagamma(h=1, z=0, size=1_000_000), repeated to fill the sample
            reps = max(1, m // 1_000_000)
            if use_ref:
                import ref

                def job(step):
                    g = np.random.Generator(np.random.PCG64(seed + step))
                    for _ in range(reps):
                        ref.random_polyagamma(1.0, 0.0, size=1_000_000, random_state=g)
            else:
                import oracle as O

                def job(step):
                    for r in range(reps):
                        O.fill2(seed, step * reps + r, np.ones(1_000_000), np.zeros(1_000_000), None)
            return job, reps * 1_000_000
        h, z = draw_operand(rng, hn, m), draw_operand(rng, zn, m)
        out = np.empty(m)
        if use_ref:
            import ref

            def job(step):
                ref.random_polyagamma(h, z, out=out, method=method, disable_checks=True,
                                      random_state=np.random.Generator(np.random.PCG64(seed + 1000 * step)))
        else:
            import oracle as O

            hh = np.broadcast_to(np.asarray(h, dtype=np.float64), (m,))

            def job(step):
                out[:] = O.fill2(seed, step, hh, z, method)
        return job, m
    kind, log = DENSITIES[workload]
    # the same broadcast pattern as the GPU arm: x along rows, (h, z) along columns
    cols = max(1, int(round(m ** 0.5)))
    rows = max(1, m // cols)
    x = np.linspace(1e-3, 12.5, rows)[:, None]
    h = rng.uniform(0.5, 25.0, cols)[None, :]
    z = rng.uniform(-10.0, 10.0, cols)[None, :]
    if use_ref:
        import ref

        f = ref.polyagamma_pdf if kind == "pdf" else ref.polyagamma_cdf

        def job(step):
            f(x, h, z, return_log=log)
    else:
        import oracle as O

        name = ("log" if log else "") + kind

        def job(step):
            O.density(name, np.broadcast_to(x, (rows, cols)).ravel(), np.broadcast_to(h, (rows, cols)).ravel(),
                      np.broadcast_to(z, (rows, cols)).ravel())
    return job, rows * cols


def _worker(idx, workload, m, start, done, times, stop, use_ref):
    job, _ = make_job(workload, m, 20240 + idx, use_ref)
    step = 0
    while True:
        start.wait()
        if stop.value:
            return
        t0 = time.perf_counter()
        job(step)
        times[idx] = time.perf_counter() - t0
        step += 1
        done.wait()


def reference_kind() -> str:
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import ref

    return "reference" if ref.available() else "port"


class ReferencePool:
    """P persistent processes; `step()` makes each do one job of m elements and returns the wall time."""

    def __init__(self, workload: str, m: int, procs: int | None = None):
        self.kind = reference_kind()
        self.procs = procs or len(os.sched_getaffinity(0))
        self.m = max(1, m // 1_000_000) * 1_000_000 if workload == "cfg1" else m
        if workload in DENSITIES:
            cols = max(1, int(round(m ** 0.5)))
            self.m = max(1, m // cols) * cols
        ctx = mp.get_context("spawn")
        self.start = ctx.Barrier(self.procs + 1)
        self.done = ctx.Barrier(self.procs + 1)
        self.times = ctx.Array("d", self.procs)
        self.stop = ctx.Value("i", 0)
        self.workers = [ctx.Process(target=_worker, args=(i, workload, m, self.start, self.done, self.times,
                                                          self.stop, self.kind == "reference"), daemon=True)
                        for i in range(self.procs)]
        for w in self.workers:
            w.start()

    def step(self, timeout: float = 900.0) -> float:
        # a broken barrier (worker died / timed out) raises instead of hanging the bench
        self.start.wait(timeout=timeout)
        t0 = time.perf_counter()
        self.done.wait(timeout=timeout)
        return time.perf_counter() - t0

    def worker_times(self):
        return list(self.times)

    def close(self):
        self.stop.value = 1
        try:
            self.start.wait(timeout=30)
        except Exception:
            pass
        for w in self.workers:
            w.join(timeout=10)
            if w.is_alive():
                w.kill()


def single_thread_rate(workload: str, m: int) -> tuple[float, str]:
    kind = reference_kind()
    job, done = make_job(workload, m, 20240, kind == "reference")
    best = 1e30
    for rep in range(2):
        t0 = time.perf_counter()
        job(rep)
        best = min(best, time.perf_counter() - t0)
    return done / best, kind

"""oracle/ref_bench.py -- TEST/BENCH INFRASTRUCTURE ONLY.

Times the reference's own CPU implementation of the hot path (pgm_random_polyagamma_fill2
through the reference's Python API, numpy PCG64) on the host cores: persistent worker
processes, one per core, each drawing a bounded sample of the bench workload; the main process
times barrier-to-barrier, so a step's time is the slowest worker's.  Used only by bench.py
(`cpu_baseline` and `--impl reference`).  Falls back to the oracle port (libpgm_oracle.so)
when oracle/_ref has not been built, and says so through `kind`.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def workload_arrays(workload: str, n: int, seed: int):
    """Synthetic inputs of a BASELINE.json config (same distributions as the GPU arm)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if workload == "cfg4":
        return rng.uniform(0.1, 200.0, n), rng.uniform(-50.0, 50.0, n), None
    if workload == "cfg2":
        return np.ones(n), rng.normal(0.0, 3.0, n), None
    raise ValueError(workload)


def _worker(idx, workload, m, start, done, times, stop, use_ref):
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    h, z, method = workload_arrays(workload, m, 20240 + idx)
    out = np.empty(m)
    if use_ref:
        import ref

        def draw(step):
            ref.random_polyagamma(h, z, out=out, method=method, disable_checks=True,
                                  random_state=np.random.Generator(np.random.PCG64(12345 + 1000 * idx + step)))
    else:
        import oracle as O

        def draw(step):
            out[:] = O.fill2(12345 + idx, step, h, z, method)
    step = 0
    while True:
        start.wait()
        if stop.value:
            return
        t0 = time.perf_counter()
        draw(step)
        times[idx] = time.perf_counter() - t0
        step += 1
        done.wait()


class ReferencePool:
    """P persistent processes; `step()` makes each draw m samples and returns the wall time."""

    def __init__(self, workload: str, m: int, procs: int | None = None):
        if HERE not in sys.path:
            sys.path.insert(0, HERE)
        import ref

        self.kind = "reference" if ref.available() else "port"
        self.procs = procs or len(os.sched_getaffinity(0))
        self.m = m
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

    def step(self, timeout: float = 600.0) -> float:
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
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import ref

    h, z, method = workload_arrays(workload, m, 20240)
    out = np.empty(m)
    if ref.available():
        best = 1e30
        for rep in range(2):
            t0 = time.perf_counter()
            ref.random_polyagamma(h, z, out=out, method=method, disable_checks=True,
                                  random_state=np.random.Generator(np.random.PCG64(rep)))
            best = min(best, time.perf_counter() - t0)
        return m / best, "reference"
    import oracle as O

    t0 = time.perf_counter()
    O.fill2(1, 0, h, z, method)
    return m / (time.perf_counter() - t0), "port"

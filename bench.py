#!/usr/bin/env python
"""bench.py -- throughput of the Polya-Gamma hot path on N B200s of one node.

    python bench.py [--config cfg4] --gpus N --steps K --warmup W      (N > 1: under torchrun, 1 rank/GPU)
    python bench.py --impl reference [--config ...] ...                (the reference on the host cores)

--config names a BASELINE.json configuration (default cfg4, the one the headline metric is quoted on):

    cfg1            random_polyagamma(h=1, z=0, size=1e6), hybrid -> devroye          (configs[0])
    cfg2            fill2 hybrid, h = ones array, z ~ N(0, 3), 1e8 elements            (configs[1])
    cfg2-scalar-h   the same with h the scalar 1.0 broadcast                           (configs[1])
    cfg3-devroye | cfg3-alternate | cfg3-saddle | cfg3-gamma   per-method sweeps, 1e8  (configs[2])
    cfg4            fill2 hybrid, h ~ U[0.1, 200], z ~ U[-50, 50], 1e9 elements        (configs[3])
    cfg5a-pdf | cfg5a-cdf | cfg5a-logpdf | cfg5a-logcdf   1e4 x 1e4 broadcast grid     (configs[4])

One step = one pass of the hot path over that batch of synthetic input.  The index range is sharded over
ranks (weak scaling: every rank owns its own consecutive global indices and passes first_index); there is no
collective on the data path.  The JSON line carries

  value         units/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e           the same metric through the public API with HOST (pinned) buffers: host->device copies of the
                inputs and the device->host copy of the result are inside the timed region
  roofline      dominant kernel (by device time, measured live per kernel with CUDA events on the launching
                stream, kernels serialised, OUTSIDE the headline timed region) against the FP64 peak measured on
                this box, plus the HBM view (algorithmic bytes) against MEASURED_PEAKS.json
  cpu_baseline  the reference (oracle/_ref) on the host cores, bounded sample (rank 0, N = 1)
  shard_checksum  64-bit sum / xor over the bit patterns of one FIXED global index range of this config drawn
                shard-wise by the N ranks (NCCL-reduced): equal for every N iff the partitioning is invisible
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEADLINE_METRIC = "PG(h,z) samples/sec at 1/2/4/8 B200 (hybrid, mixed h,z); % of roofline"

# name -> (kind, default elements per GPU and step, description)
CONFIGS = {
    "cfg1": ("fill", 1_000_000, "cfg1: random_polyagamma(h=1, z=0, size=1e6) hybrid -> devroye (BASELINE.json configs[0])"),
    "cfg2": ("fill2", 100_000_000, "cfg2: fill2 hybrid, h=ones array, z~N(0,3) (BASELINE.json configs[1])"),
    "cfg2-scalar-h": ("fill2", 100_000_000, "cfg2: fill2 hybrid, h=1.0 scalar broadcast, z~N(0,3) (BASELINE.json configs[1])"),
    "cfg3-devroye": ("fill2", 100_000_000, "cfg3: fill2 devroye, h=1, z~U[-50,50] (BASELINE.json configs[2])"),
    "cfg3-alternate": ("fill2", 100_000_000, "cfg3: fill2 alternate, h~U[1,4], z~U[-50,50] (BASELINE.json configs[2])"),
    "cfg3-saddle": ("fill2", 100_000_000, "cfg3: fill2 saddle, h~U[10,100], z~U[-50,50] (BASELINE.json configs[2])"),
    "cfg3-gamma": ("fill2", 100_000_000, "cfg3: fill2 gamma, h~U[0.1,1], z~U[-50,50] (BASELINE.json configs[2])"),
    "cfg4": ("fill2", 1_000_000_000, "cfg4: fill2 hybrid, h~U[0.1,200], z~U[-50,50] (BASELINE.json configs[3])"),
    "cfg5a-pdf": ("density", 100_000_000, "cfg5a: polyagamma_pdf on a 1e4 x 1e4 broadcast grid (BASELINE.json configs[4])"),
    "cfg5a-cdf": ("density", 100_000_000, "cfg5a: polyagamma_cdf on a 1e4 x 1e4 broadcast grid (BASELINE.json configs[4])"),
    "cfg5a-logpdf": ("density", 100_000_000, "cfg5a: polyagamma_pdf(return_log=True) on a 1e4 x 1e4 broadcast grid (BASELINE.json configs[4])"),
    "cfg5a-logcdf": ("density", 100_000_000, "cfg5a: polyagamma_cdf(return_log=True) on a 1e4 x 1e4 broadcast grid (BASELINE.json configs[4])"),
}
# workload -> (h, z, method) as in oracle/ref_bench.py (the CPU arm draws the same distributions)
SAMPLERS = {
    "cfg1": ("one", "zero", None), "cfg2": ("ones", "normal3", None), "cfg2-scalar-h": ("one", "normal3", None),
    "cfg3-devroye": ("ones", "u50", "devroye"), "cfg3-alternate": ("u1_4", "u50", "alternate"),
    "cfg3-saddle": ("u10_100", "u50", "saddle"), "cfg3-gamma": ("u01_1", "u50", "gamma"), "cfg4": ("u01_200", "u50", None),
}
DENSITIES = {"cfg5a-pdf": ("pdf", False), "cfg5a-logpdf": ("pdf", True), "cfg5a-cdf": ("cdf", False),
             "cfg5a-logcdf": ("cdf", True)}
# kernel kind (polyagamma_b200._lib.KERNEL_KINDS) -> kernel name
KERNEL_NAMES = {"tile": "tile_kernel (classify + in-tile normal / left-saddle + list compaction)",
                "small": "small_kernel", "devroye": "sample_kernel<DevroyeT<pair|MSH>>",
                "alternate": "sample_kernel<Alternate>", "alternate_left": "sample_kernel<AlternateLeft, fused setup>",
                "saddle": "sample_kernel<SaddleFull>", "saddle_left": "sample_kernel<SaddleLeft>",
                "normal": "dense_kernel<normal>", "gamma": "dense_kernel<gamma>", "density": "density_kernel",
                "setup_devroye": "setup_kernel<DevroyeT<pair|MSH>>", "setup_alternate": "setup_kernel<Alternate>",
                "setup_saddle": "setup_kernel<SaddleFull>", "setup_saddle_left": "setup_kernel<SaddleLeft>"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=float, default=None, help="elements per GPU per step (default: the config's size)")
    ap.add_argument("--e2e-elems", type=float, default=None, help="elements per GPU per e2e step")
    ap.add_argument("--ref-elems", type=float, default=None, help="reference arm: elements per core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-checksum", action="store_true")
    return ap.parse_args()


def metric_of(config):
    if config == "cfg4":
        return HEADLINE_METRIC, "samples/s"
    if config in DENSITIES:
        return f"PG(h,z) density points/sec ({config}); % of roofline", "points/s"
    return f"PG(h,z) samples/sec ({config}); % of roofline", "samples/s"


# --------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def _nvml_handle(self):
        """NVML handle of this rank's GPU (by the UUID torch reports: CUDA_VISIBLE_DEVICES may renumber), or None."""
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.gpu).uuid)
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except TypeError:
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            return pynvml, h
        except Exception:
            return None

    def _poll_nvml(self, pynvml, h):
        names = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        try:
            smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            smax = None
        while not self.stop_flag.is_set():
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = int(get_reasons(h))
                try:
                    power = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    power = None
                self.nvml_samples.append((sm, smax, power, [k for k, bit in names.items() if mask & bit]))
            except Exception:
                pass
            self.stop_flag.wait(0.004)

    def start(self):
        self.nvml_samples = []
        self.stop_flag = threading.Event()
        got = self._nvml_handle()
        if got is not None:
            self.thread = threading.Thread(target=self._poll_nvml, args=got, daemon=True)
            self.thread.start()
            self.mode = "nvml"
            return
        self.mode = "nvidia-smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if getattr(self, "mode", "") == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=2)
            sm = [x[0] for x in self.nvml_samples]
            smax = [x[1] for x in self.nvml_samples if x[1] is not None]
            power = [x[2] for x in self.nvml_samples if x[2] is not None]
            reasons = sorted({r for x in self.nvml_samples for r in x[3]})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                    "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": reasons,
                    "source": "NVML, polled every 4 ms during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "MEASURED_PEAKS.json (of measured)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "B200_PROFILING.md fallback (of fallback)"


def kernel_costs(config):
    """Executed FP64 flops and DRAM bytes per unit of each kernel on this config's distribution, from the ncu
    capture summarised in profiles/r2_kernel_costs.json (tools/ncu_costs.py)."""
    path = os.path.join(ROOT, "profiles", "r2_kernel_costs.json")
    try:
        return json.load(open(path)).get(config, {})
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# --------------------------------------------------------------------------------------------
def reference_arm(args, rank):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_bench

    metric, unit = metric_of(args.config)
    m = int(args.ref_elems) if args.ref_elems else ref_bench.DEFAULT_SAMPLE[args.config]
    pool = ref_bench.ReferencePool(args.config, m)
    try:
        for _ in range(args.warmup):
            pool.step()
        t = [pool.step() for _ in range(args.steps)]
    finally:
        pool.close()
    total = sum(t)
    value = pool.procs * pool.m * args.steps / total
    sample = f"{pool.procs} processes x {pool.m} {args.config} elements per step, {args.steps} steps"
    emit({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config][2], "elems_per_core_per_step": pool.m, "host_cores": pool.procs,
                   "rng": "numpy PCG64", "call": "the reference's Python API (random_polyagamma(h, z, out=, "
                   "disable_checks=True) -> pgm_random_polyagamma_fill2; polyagamma_pdf / polyagamma_cdf -> dispatch)"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": pool.procs, "kind": pool.kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def cpu_baseline(config, unit):
    """Bounded all-core + single-thread run of the reference on this box's host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_bench

    m = ref_bench.DEFAULT_SAMPLE[config]
    single, kind = ref_bench.single_thread_rate(config, m)
    pool = ref_bench.ReferencePool(config, m)
    try:
        pool.step()
        t = min(pool.step() for _ in range(3))
    finally:
        pool.close()
    return {"value": pool.procs * pool.m / t, "unit": unit, "cores": pool.procs, "kind": pool.kind,
            "single_thread": single,
            "sample": f"{pool.procs} processes x {pool.m} {config} elements (best of 3); single thread {pool.m} "
                      "elements (best of 2)"}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# Libraries chat on stdout (NCCL prints its version banner there): keep the real stdout for the
# JSON line only and send everything else to stderr.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def device_operand(torch, name, n, dev, gen):
    """Synthetic operand on the device (same distributions as oracle/ref_bench.draw_operand)."""
    f64 = torch.float64
    if name == "one":
        return 1.0
    if name == "zero":
        return 0.0
    if name == "ones":
        return torch.ones(n, device=dev, dtype=f64)
    if name == "normal3":
        return torch.randn(n, device=dev, dtype=f64, generator=gen) * 3.0
    if name == "u50":
        return torch.rand(n, device=dev, dtype=f64, generator=gen) * 100.0 - 50.0
    lo, hi = {"u1_4": (1.0, 4.0), "u10_100": (10.0, 100.0), "u01_1": (0.1, 1.0), "u01_200": (0.1, 200.0)}[name]
    return torch.rand(n, device=dev, dtype=f64, generator=gen) * (hi - lo) + lo


class Workload:
    """Device-resident and host-buffer (e2e) forms of one config."""

    def __init__(self, config, n, rank, dev, torch, pg):
        self.config, self.n, self.rank, self.dev, self.torch, self.pg = config, n, rank, dev, torch, pg
        self.kind = CONFIGS[config][0]
        self.first_index = rank * n
        gen = torch.Generator(device=dev)
        gen.manual_seed(20240 + rank)
        f64 = torch.float64
        if self.kind == "density":
            self.fname, self.log = DENSITIES[config]
            self.rows = self.cols = int(round(n ** 0.5))
            self.n = self.rows * self.cols
            self.x = torch.linspace(1e-3, 12.5, self.rows, device=dev, dtype=f64)[:, None]
            self.h = (torch.rand(self.cols, device=dev, dtype=f64, generator=gen) * 24.5 + 0.5)[None, :]
            self.z = (torch.rand(self.cols, device=dev, dtype=f64, generator=gen) * 20.0 - 10.0)[None, :]
            self.alg_bytes = 8.0  # one result per point; the broadcast operands are 24 B per row / column
            self.result = None
        else:
            hn, zn, self.method = SAMPLERS[config]
            self.h = device_operand(torch, hn, n, dev, gen)
            self.z = device_operand(torch, zn, n, dev, gen)
            self.out = torch.empty(n, device=dev, dtype=f64)
            self.alg_bytes = 8.0 + (8.0 if torch.is_tensor(self.h) else 0.0) + (8.0 if torch.is_tensor(self.z) else 0.0)
        self.footprint = self.alg_bytes * self.n

    # ---- device resident ----
    def step(self, i):
        pg = self.pg
        if self.kind == "density":
            f = pg.polyagamma_pdf if self.fname == "pdf" else pg.polyagamma_cdf
            self.result = f(self.x, self.h, self.z, return_log=self.log)
        elif self.kind == "fill":
            self.out = pg.random_polyagamma(self.h, self.z, size=self.n, device=self.dev, random_state=(12345, i),
                                            first_index=self.first_index)
        else:
            pg.random_polyagamma(self.h, self.z, out=self.out, method=self.method, random_state=(12345, i),
                                 disable_checks=True, first_index=self.first_index)

    def last_output(self):
        return self.result if self.kind == "density" else self.out

    # ---- host buffers ----
    def host_setup(self, ne):
        torch = self.torch
        f64 = torch.float64

        def pin(t):
            hbuf = torch.empty(t.shape, dtype=f64, pin_memory=True)
            hbuf.copy_(t)
            return hbuf.numpy()

        self.ne = ne
        if self.kind == "density":
            r = max(1, min(self.rows, ne // self.cols))
            self.ne = r * self.cols
            self.hx, self.hh, self.hz = pin(self.x[:r]), pin(self.h), pin(self.z)
            self.h2d, self.d2h = 8 * (r + 2 * self.cols), 8 * self.ne
        else:
            self.hh = pin(self.h[:ne]) if torch.is_tensor(self.h) else self.h
            self.hz = pin(self.z[:ne]) if torch.is_tensor(self.z) else self.z
            self.ho_t = torch.empty(ne, dtype=f64, pin_memory=True)
            self.ho = self.ho_t.numpy()
            self.h2d = 8 * ne * (int(torch.is_tensor(self.h)) + int(torch.is_tensor(self.z)))
            self.d2h = 8 * ne
        torch.cuda.synchronize()

    def host_step(self, i):
        pg = self.pg
        if self.kind == "density":
            f = pg.polyagamma_pdf if self.fname == "pdf" else pg.polyagamma_cdf
            self.hres = f(self.hx, self.hh, self.hz, return_log=self.log)
        elif self.kind == "fill":
            pg.random_polyagamma(self.hh, self.hz, out=self.ho, random_state=(777, i), first_index=self.first_index)
        else:
            pg.random_polyagamma(self.hh, self.hz, out=self.ho, method=self.method, random_state=(777, i),
                                 disable_checks=True, first_index=self.first_index)

    def host_result_finite(self):
        import numpy as np

        arr = self.hres if self.kind == "density" else self.ho
        flat = arr.reshape(-1)
        v = flat[:: max(1, flat.size // 1_000_000)]
        return bool(np.all(np.isfinite(v) | (v == -np.inf))) if self.kind == "density" else bool(np.all(np.isfinite(v)))

    def api(self):
        if self.kind == "density":
            return f"polyagamma_b200.polyagamma_{self.fname}(x_host[:, None], h_host[None, :], z_host[None, :]" + \
                   (", return_log=True)" if self.log else ")") + " -> numpy result"
        return ("polyagamma_b200.random_polyagamma(h_host, z_host, out=out_host) -> "
                "pgm_cuda_random_polyagamma_fill2_host (pinned buffers)")


def host_link_peak(torch, dev, barrier, max_over_ranks):
    """GB/s per GPU of concurrent pinned H2D (2 parts) + D2H (1 part) copies on three streams, all ranks at once."""
    ch = 1 << 23
    hin = [torch.empty(2 * ch, dtype=torch.float64, pin_memory=True).fill_(1.0) for _ in range(3)]
    hout = [torch.empty(ch, dtype=torch.float64, pin_memory=True) for _ in range(3)]
    din = [torch.empty(2 * ch, dtype=torch.float64, device=dev) for _ in range(3)]
    dout = [torch.zeros(ch, dtype=torch.float64, device=dev) for _ in range(3)]
    streams = [torch.cuda.Stream(dev) for _ in range(3)]
    nch, best = 9, None
    for rep in range(3):
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for c in range(nch):
            with torch.cuda.stream(streams[c % 3]):
                din[c % 3].copy_(hin[c % 3], non_blocking=True)
                hout[c % 3].copy_(dout[c % 3], non_blocking=True)
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3
        if rep and (best is None or dt < best):
            best = dt
    return 24 * ch * nch / best / 1e9


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric, unit = metric_of(args.config)

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    # the CPU baseline forks/spawns processes: run it before CUDA is initialised
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(args.config, unit)
        except Exception as exc:  # the number is a reported baseline, never a reason to fail
            cpu = {"value": None, "unit": unit, "cores": 0, "kind": "unavailable", "sample": repr(exc)}

    import torch
    import torch.distributed as dist

    import polyagamma_b200 as pg
    from polyagamma_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = int(args.elems) if args.elems else CONFIGS[args.config][1]
    wl = Workload(args.config, n, rank, dev, torch, pg)
    n = wl.n
    fp64_peak_measured = _lib.fp64_peak_tflops()

    # Inputs + output smaller than twice the L2 are flushed out of it between timed steps (a 512 MB write);
    # larger ones stream through it anyway.
    flush = None
    if wl.footprint < 2 * 126e6 * 2:
        flush = torch.empty(512 * 1024 * 1024 // 8, device=dev, dtype=torch.float64)

    for i in range(args.warmup):
        wl.step(i)
    torch.cuda.synchronize()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = pg.launch_count()
    torch.cuda.synchronize()
    if flush is None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            wl.step(args.warmup + i)
        e1.record()
        torch.cuda.synchronize()
        elapsed_local = e0.elapsed_time(e1)
    else:
        pairs = []
        for i in range(args.steps):
            flush.fill_(float(i))
            ea = torch.cuda.Event(enable_timing=True)
            eb = torch.cuda.Event(enable_timing=True)
            ea.record()
            wl.step(args.warmup + i)
            eb.record()
            pairs.append((ea, eb))
        torch.cuda.synchronize()
        elapsed_local = sum(a.elapsed_time(b) for a, b in pairs)
    launches = pg.launch_count() - launches0
    barrier()
    elapsed_ms = max_over_ranks(elapsed_local)
    clock_info = clocks.stop()
    value = world * n * args.steps / (elapsed_ms * 1e-3)
    outv = wl.last_output().reshape(-1)
    sample = outv[:: max(1, outv.numel() // 1_000_000)]
    finite = bool((torch.isfinite(sample) | (sample == float("-inf"))).all()) if wl.kind == "density" else \
        bool(torch.isfinite(sample).all())

    # ---- per-kernel times: the same K steps once more, profiled, with the kernels serialised on one stream
    # (in the timed region above the buckets of a pass overlap on helper streams, and nothing is profiled) ----
    os.environ["PGM_CUDA_NO_OVERLAP"] = "1"
    wl.step(0)
    torch.cuda.synchronize()
    _lib.profile_begin()
    for i in range(args.steps):
        wl.step(args.warmup + i)
    torch.cuda.synchronize()
    kernels = _lib.profile_end()
    del os.environ["PGM_CUDA_NO_OVERLAP"]

    # ---- roofline of the dominant kernel ----
    peaks, peaks_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    fp64_nominal = 148 * 64 * 2 * sm_max_mhz * 1e6 / 1e12  # 148 SMs x 64 DFMA/clk x 2 flops
    fp64_peak = max(fp64_peak_measured, fp64_nominal)  # the larger of the two: never flatter the fraction
    costs = kernel_costs(args.config)
    active = {k: v for k, v in kernels.items() if v[1] and v[0] > 0}
    total_kernel_ms = sum(v[0] for v in active.values())
    dom = max(active, key=lambda k: active[k][0]) if active else None
    roofline = None
    if dom is not None:
        dom_ms, dom_launches = active[dom]
        launch_s = dom_ms * 1e-3 / dom_launches
        units_per_launch = n * args.steps / dom_launches  # (the dominant kernel sees every unit of the step)
        c = costs.get(dom, {})
        hbm_achieved = units_per_launch * wl.alg_bytes / launch_s / 1e9
        fp64 = None
        if c.get("flops_per_unit"):
            tf = units_per_launch * c["flops_per_unit"] / launch_s / 1e12
            fp64 = {"achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                    "flops_per_unit": c["flops_per_unit"],
                    "flops_source": "executed 2*DFMA + DMUL + DADD thread instructions of one ncu-profiled launch / "
                                    "its units (profiles/r2_kernel_costs.json)"}
        hbm = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
               "bytes_per_unit": wl.alg_bytes, "peak_source": peaks_src}
        lead = fp64 if (fp64 and fp64["frac"] >= hbm["frac"]) else hbm
        roofline = {
            "bound": "fp64" if lead is fp64 else "hbm", "kernel": KERNEL_NAMES.get(dom, dom),
            "achieved": lead["achieved"], "peak": lead["peak"], "unit": lead["unit"], "frac": lead["frac"],
            "traffic": units_per_launch * c["dram_bytes_per_unit"] if c.get("dram_bytes_per_unit") else None,
            "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per unit (profiles/r2_kernel_costs.json) "
                              "x units of this launch",
            "algorithmic_bytes_per_launch": units_per_launch * wl.alg_bytes,
            "units_per_launch": units_per_launch, "launch_ms": launch_s * 1e3,
            "share_of_step": dom_ms / total_kernel_ms if total_kernel_ms else None,
            "fp64": fp64, "hbm": hbm,
            "fp64_peak": {"measured_tflops": fp64_peak_measured, "nominal_tflops": fp64_nominal,
                          "note": "measured: 16-chain DFMA microbenchmark in this process (builder-measured: "
                                  "MEASURED_PEAKS.json has no FP64 entry); nominal: 148 SMs x 64 DFMA/clk x 2 x max SM clock "
                                  "(tools/dmma_peak.cu reaches 37.0 with DMMA); fractions use the LARGER of the two"},
            "timing": "per-kernel times: the same K steps run once more with CUDA events around every launch and "
                      "the kernels serialised (PGM_CUDA_NO_OVERLAP=1), outside the headline timed region",
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in active.items()},
            "kernel_launches_per_step": {k: v[1] / args.steps for k, v in active.items()},
            "whole_step": {"hbm_gbs": n * args.steps * wl.alg_bytes / (elapsed_local * 1e-3) / 1e9,
                           "hbm_frac": n * args.steps * wl.alg_bytes / (elapsed_local * 1e-3) / 1e9 / hbm_peak},
        }
        if costs:
            tot_flops = sum(n * c2.get("share_of_units", 1.0) * c2.get("flops_per_unit", 0.0) for k2, c2 in costs.items()
                            if k2 in active)
            roofline["whole_step"].update({"fp64_tflops": tot_flops * args.steps / (elapsed_local * 1e-3) / 1e12,
                                           "fp64_frac": tot_flops * args.steps / (elapsed_local * 1e-3) / 1e12 / fp64_peak})

    # ---- one fixed global index range drawn shard-wise: the partitioning must be invisible ----
    checksum = None
    if wl.kind != "density" and not args.no_checksum:
        G = 1 << 26 if args.config != "cfg3-gamma" else 1 << 20
        gen = torch.Generator(device=dev)
        gen.manual_seed(4242)
        hn, zn, method = SAMPLERS[args.config]
        hg, zg = device_operand(torch, hn, G, dev, gen), device_operand(torch, zn, G, dev, gen)
        lo, hi = rank * G // world, (rank + 1) * G // world
        hs = hg[lo:hi] if torch.is_tensor(hg) else hg
        zs = zg[lo:hi] if torch.is_tensor(zg) else zg
        if torch.is_tensor(hs) or torch.is_tensor(zs):
            shard = pg.random_polyagamma(hs, zs, method=method, random_state=(2024, 7), disable_checks=True,
                                         first_index=lo, device=dev)
        else:
            shard = pg.random_polyagamma(hs, zs, size=hi - lo, random_state=(2024, 7), first_index=lo, device=dev)
        bits = shard.reshape(-1).view(torch.int64)
        ssum = bits.sum()  # wraps modulo 2^64
        sxor = torch.zeros((), dtype=torch.int64, device=dev)
        for chunk in bits.split(1 << 24):
            # xor-reduce by halving (torch has no xor reduction)
            c = chunk
            while c.numel() > 1:
                if c.numel() & 1:
                    c = torch.cat([c, torch.zeros(1, dtype=torch.int64, device=dev)])
                half = c.numel() // 2
                c = c[:half] ^ c[half:]
            sxor ^= c[0]
        if world > 1:
            dist.all_reduce(ssum, op=dist.ReduceOp.SUM)
            parts = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(parts, sxor)
            sxor = parts[0]
            for p in parts[1:]:
                sxor = sxor ^ p
        checksum = {"global_elems": G, "seed": 2024, "offset": 7, "shards": world,
                    "sum_u64": f"0x{int(ssum.item()) & ((1 << 64) - 1):016x}",
                    "xor_u64": f"0x{int(sxor.item()) & ((1 << 64) - 1):016x}",
                    "note": "bit patterns of the draws of global indices [0, G) with this config's distribution, drawn by "
                            "the N ranks shard-wise with first_index and reduced over NCCL: identical for every N"}
        del hg, zg, shard, bits

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        import psutil

        want = int(args.e2e_elems) if args.e2e_elems else n
        budget = int(psutil.virtual_memory().available * 0.4 / max(world, 1) / 24)
        ne = max(1 << 16, min(want, budget, n))
        wl.host_setup(ne)
        for i in range(max(args.warmup, 1)):
            wl.host_step(i)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            wl.host_step(100 + i)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        # every call blocks until its host result is complete: the host clock around the K calls is the measure
        e2e_ms = max_over_ranks(wall_ms)
        e2e = {"value": world * wl.ne * args.steps / (e2e_ms * 1e-3), "unit": unit, "h2d_bytes_per_step": wl.h2d,
               "d2h_bytes_per_step": wl.d2h, "elems_per_gpu": wl.ne, "ms_per_step": e2e_ms / args.steps,
               "api": wl.api(), "result_finite": wl.host_result_finite()}
        if wl.kind != "density":
            link = (wl.h2d + wl.d2h) * args.steps / (e2e_ms * 1e-3) / 1e9
            e2e["host_link_gbs"] = link
            # the roof of this number is the host link, and boxes differ: measure it here and now with plain
            # copies in the call's own pattern (16 B in + 8 B out per element, three streams, every rank at once)
            live = host_link_peak(torch, dev, barrier, max_over_ranks)
            recorded = None
            try:
                hp = json.load(open(os.path.join(ROOT, "profiles", "r2_h2d_peak.json")))
                recorded = hp.get(str(world), {}).get("bidir_gbs_per_gpu")
            except Exception:
                pass
            e2e["roofline"] = {"achieved_gbs": link, "peak_gbs": live, "frac": link / live,
                               "peak_source": "concurrent pinned H2D + D2H per GPU at this N, measured in this run after "
                                              "the timed region (the loop of tools/h2d_peak.py)",
                               "recorded_peak_gbs": recorded,
                               "recorded_peak_source": "profiles/r2_h2d_peak.json (another box of the pool)"}

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config][2], "name": args.config, "elems_per_gpu": n,
                       "global_elems": world * n, "parallelism": f"index-range sharding x{world}, no collective",
                       "l2": ("a 512 MB buffer is written between timed steps (working set below 2x the 126 MB L2); "
                              "each step is timed by its own CUDA-event pair" if flush is not None else
                              f"working set {wl.footprint / 1e9:.1f} GB per step per GPU (>> 126 MB L2): no flush needed"),
                       "seed": 12345},
            "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "result_finite": finite, "shard_checksum": checksum,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

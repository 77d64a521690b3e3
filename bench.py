#!/usr/bin/env python
"""bench.py -- PG(h, z) samples/s of the hybrid batched fill, on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (N > 1: under torchrun, 1 rank/GPU)
    python bench.py --impl reference ...                     (the reference on the host cores)

Workload = BASELINE.json configs[3] ("cfg4", the configuration the metric is quoted on):
fill2 with the hybrid method over h ~ U[0.1, 200], z ~ U[-50, 50], 1e9 elements per GPU,
synthetic inputs.  One step = one pass of the hot path over that batch.  The index range is
sharded over ranks (weak scaling: every rank owns 1e9 consecutive global indices and passes
first_index); there is no collective on the data path.

The JSON line carries
  value      samples/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with HOST (pinned) buffers: host->device copies
             of h and z and the device->host copy of the result are inside the timed region
  roofline   dominant kernel (by device time, measured live per kernel with CUDA events on the
             launching stream) against the FP64 DFMA peak measured on this box, plus the HBM
             view against MEASURED_PEAKS.json
  cpu_baseline  the reference (oracle/_ref) on the host cores, bounded sample (rank 0, N = 1)
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

METRIC = "PG(h,z) samples/sec at 1/2/4/8 B200 (hybrid, mixed h,z); % of roofline"
UNIT = "samples/s"
WORKLOAD = "cfg4: fill2 hybrid, h~U[0.1,200], z~U[-50,50] (BASELINE.json configs[3])"

# FP64 flops per element of each kernel on the cfg4 distribution: executed thread-level
# instructions 2*DFMA + DMUL + DADD (ncu smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on)
# of one profiled launch divided by the elements of that launch -- profiles/r1j_cfg4_all_kernels.txt,
# table in profiles/README.md (SURVEY 8(d)'s nominal per-method figures are kept beside them there).
# devroye figures are from the cfg2 capture (cfg4 has no devroye elements; its setup is fused as well).
FLOPS_PER_ELEMENT = {"devroye": 254.0, "alternate": 508.0, "alternate_left": 384.0, "saddle": 600.0,
                     "saddle_left": 640.0, "saddle_far": 577.0, "saddle_mid": 645.0, "normal": 152.0, "gamma": 6.0e4,
                     "setup_devroye": 0.0, "setup_alternate": 1006.0, "setup_saddle": 1569.0,
                     "setup_saddle_left": 354.0,
                     # the far / mid saddle and alternate-left buckets compute their records inside the
                     # sampling kernel: no setup launch, the sampling figure includes the setup
                     "setup_alternate_left": 0.0, "setup_saddle_far": 0.0, "setup_saddle_mid": 0.0}
# DRAM bytes per element of each kernel: (dram__bytes_read.sum + dram__bytes_write.sum) of the same ncu
# capture divided by the elements of that launch.  Algorithmic bytes are 28 (normal: list index + h + z +
# out) to ~90 (record kernels); the sparse buckets pay 32-byte sectors for 8-byte gathers of h and z.
DRAM_BYTES_PER_ELEMENT = {"normal": 42.0, "saddle_far": 146.9, "saddle_mid": 270.6, "saddle_left": 104.3,
                          "saddle": 115.8, "alternate": 75.3, "alternate_left": 275.6,
                          "setup_saddle_left": 241.2, "setup_saddle": 256.4, "setup_alternate": 253.8}
KERNEL_NAMES = {"devroye": "sample_kernel<DevroyeT<pair|MSH>>", "alternate": "sample_kernel<Alternate>",
                "alternate_left": "sample_kernel<AlternateLeft, fused setup>", "setup_alternate_left": "setup_kernel<AlternateLeft>",
                "saddle": "sample_kernel<SaddleFull>", "saddle_left": "sample_kernel<SaddleLeft>",
                "saddle_far": "sample_kernel<SaddleFar, fused setup>", "saddle_mid": "sample_kernel<SaddleFar, fused setup> (8<=|z|<14)",
                "setup_saddle_mid": "setup_kernel<SaddleFar> (8<=|z|<14)",
                "normal": "dense_kernel<normal>", "gamma": "dense_kernel<gamma>",
                "setup_devroye": "setup_kernel<DevroyeT<pair|MSH>>", "setup_alternate": "setup_kernel<Alternate>",
                "setup_saddle": "setup_kernel<SaddleFull>", "setup_saddle_left": "setup_kernel<SaddleLeft>",
                "setup_saddle_far": "setup_kernel<SaddleFar>"}
BYTES_PER_ELEMENT = 24 + 4  # h, z in + out + the 4-byte list index of the bucketing pass


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=float, default=1e9, help="elements per GPU per step")
    ap.add_argument("--e2e-elems", type=float, default=None, help="elements per GPU per e2e step")
    ap.add_argument("--ref-elems", type=float, default=2e6, help="reference arm: elements per core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


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

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
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


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# --------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_bench

    m = int(args.ref_elems)
    pool = ref_bench.ReferencePool("cfg4", m)
    try:
        for _ in range(args.warmup):
            pool.step()
        t = [pool.step() for _ in range(args.steps)]
    finally:
        pool.close()
    total = sum(t)
    value = pool.procs * m * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elems_per_core_per_step": m, "host_cores": pool.procs,
                   "rng": "numpy PCG64", "call": "random_polyagamma(h, z, out=out, disable_checks=True) -> "
                   "pgm_random_polyagamma_fill2"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": pool.procs, "kind": pool.kind,
                         "sample": f"{pool.procs} processes x {m} cfg4 elements per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline():
    """Bounded all-core + single-thread run of the reference on this box's host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_bench

    m = 2_000_000
    single, kind = ref_bench.single_thread_rate("cfg4", m)
    pool = ref_bench.ReferencePool("cfg4", m)
    try:
        pool.step()
        t = min(pool.step() for _ in range(3))
    finally:
        pool.close()
    return {"value": pool.procs * m / t, "unit": UNIT, "cores": pool.procs, "kind": pool.kind,
            "single_thread": single,
            "sample": f"{pool.procs} processes x {m} cfg4 elements (best of 3); single thread {m} elements (best of 2)"}


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


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    # the CPU baseline forks/spawns processes: run it before CUDA is initialised
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline()
        except Exception as exc:  # the number is a reported baseline, never a reason to fail
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(exc)}

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

    n = int(args.elems)
    first_index = rank * n
    g = torch.Generator(device=dev)
    g.manual_seed(20240 + rank)
    h = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 199.9 + 0.1
    z = torch.rand(n, device=dev, dtype=torch.float64, generator=g) * 100.0 - 50.0
    out = torch.empty(n, device=dev, dtype=torch.float64)

    fp64_peak = _lib.fp64_peak_tflops()

    def step(i):
        pg.random_polyagamma(h, z, out=out, random_state=(12345, i), disable_checks=True, first_index=first_index)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = pg.launch_count()
    _lib.profile_begin()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    kernels_overlapped = _lib.profile_end()
    launches = pg.launch_count() - launches0
    clock_info = clocks.stop()
    # In the timed region the kernels of a pass run concurrently (helper streams), so a kernel's
    # event-bracketed time includes the others' share of the SMs.  Per-kernel roofline numbers
    # come from the same K steps run once more with the kernels serialised on one stream.
    os.environ["PGM_CUDA_NO_OVERLAP"] = "1"
    step(0)
    torch.cuda.synchronize()
    _lib.profile_begin()
    for i in range(args.steps):
        step(args.warmup + i)
    torch.cuda.synchronize()
    kernels = _lib.profile_end()
    del os.environ["PGM_CUDA_NO_OVERLAP"]
    value = world * n * args.steps / (elapsed_ms * 1e-3)
    finite = bool(torch.isfinite(out[:: max(1, n // 1_000_000)]).all())

    # ---- roofline of the dominant kernel ----
    peaks, peaks_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    normal = h > 50
    saddle = ~normal & ((h >= 8) | ((h > 4) & (z <= 4)))
    devroye = ~normal & ~saddle & ((h == 1) | ((h == torch.floor(h)) & (z <= 1)))
    mix = {"normal": int(normal.sum()), "saddle": int(saddle.sum()), "devroye": int(devroye.sum())}
    mix["alternate"] = n - sum(mix.values())
    left_mask = saddle & (h * (1.0 + 0.125 * z.abs()) >= 22.0)
    saddle_far_count = int((left_mask & (z.abs() >= 14.0)).sum())
    saddle_mid_count = int((left_mask & (z.abs() >= 8.0)).sum()) - saddle_far_count
    saddle_left_count = int(left_mask.sum()) - saddle_far_count - saddle_mid_count
    alt_left_count = int((~normal & ~saddle & ~devroye & (z.abs() >= 14.0)).sum())
    del left_mask
    del normal, saddle, devroye
    # saddle elements whose right envelope piece is unreachable run the left-only kernels
    left, far, mid = saddle_left_count, saddle_far_count, saddle_mid_count
    units = {"devroye": mix["devroye"], "alternate": mix["alternate"] - alt_left_count,
             "alternate_left": alt_left_count, "saddle": mix["saddle"] - left - far - mid,
             "saddle_left": left, "saddle_far": far, "saddle_mid": mid, "normal": mix["normal"], "gamma": 0}
    units.update({"setup_" + k: units[k] for k in ("devroye", "alternate", "alternate_left", "saddle", "saddle_left",
                                                    "saddle_far", "saddle_mid")})
    dom = max(units, key=lambda k: kernels[k][0])
    dom_ms, dom_launches = kernels[dom]
    per_launch_s = dom_ms * 1e-3 / max(dom_launches, 1)
    units_per_launch = units[dom] * args.steps / max(dom_launches, 1)
    achieved = units_per_launch * FLOPS_PER_ELEMENT[dom] / per_launch_s / 1e12
    total_kernel_ms = sum(v[0] for v in kernels.values())
    step_flops = sum(units[k] * FLOPS_PER_ELEMENT[k] for k in units if kernels[k][1])
    aggregate_tflops = step_flops * args.steps / (elapsed_ms * 1e-3) / 1e12
    roofline = {
        "bound": "fp64", "kernel": KERNEL_NAMES[dom],
        "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
        "peak_source": "DFMA microbenchmark (pgm_cuda_fp64_peak_tflops) run in this process",
        "flops_per_unit": FLOPS_PER_ELEMENT[dom], "units_per_launch": units_per_launch,
        "launch_ms": per_launch_s * 1e3, "share_of_step": dom_ms / total_kernel_ms if total_kernel_ms else None,
        "traffic": (units_per_launch * DRAM_BYTES_PER_ELEMENT[dom]) if dom in DRAM_BYTES_PER_ELEMENT else None,
        "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per element (profiles/r1j_cfg4_all_kernels.txt) "
                          "x elements of this launch",
        "timing": "per-kernel times: same K steps with the kernels serialised (PGM_CUDA_NO_OVERLAP=1); the timed "
                  "region itself overlaps the kernels of a pass on helper streams",
        "whole_step": {"fp64_flops_per_step": step_flops, "tflops": aggregate_tflops,
                       "frac": aggregate_tflops / fp64_peak,
                       "note": "all kernels' executed FP64 flops / timed step, against the same DFMA peak"},
        "hbm": {"achieved_gbs": n * args.steps * BYTES_PER_ELEMENT / (elapsed_ms * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "frac": n * args.steps * BYTES_PER_ELEMENT / (elapsed_ms * 1e-3) / 1e9 / hbm_peak,
                "bytes_per_unit": BYTES_PER_ELEMENT, "peak_source": peaks_src},
        "per_kernel": {k: {"ms_per_step": kernels[k][0] / args.steps, "units_per_step": units[k],
                           "tflops": units[k] * args.steps * FLOPS_PER_ELEMENT[k] / (kernels[k][0] * 1e-3) / 1e12,
                           "frac": units[k] * args.steps * FLOPS_PER_ELEMENT[k] / (kernels[k][0] * 1e-3) / 1e12 / fp64_peak}
                       for k in units if kernels[k][1] and kernels[k][0] > 0 and units[k] > 0},
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kernels.items() if v[1]},
        "kernel_ms_per_step_overlapped": {k: v[0] / args.steps for k, v in kernels_overlapped.items() if v[1]},
        "kernel_launches": {k: v[1] for k, v in kernels.items() if v[1]},
    }

    # ---- end to end through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        import psutil

        want = int(args.e2e_elems) if args.e2e_elems else n
        budget = int(psutil.virtual_memory().available * 0.4 / max(world, 1) / 24)
        ne = max(1 << 20, min(want, budget, n))
        hh = torch.empty(ne, dtype=torch.float64, pin_memory=True)
        zh = torch.empty(ne, dtype=torch.float64, pin_memory=True)
        oh = torch.empty(ne, dtype=torch.float64, pin_memory=True)
        hh.copy_(h[:ne])
        zh.copy_(z[:ne])
        torch.cuda.synchronize()
        hn, zn, on = hh.numpy(), zh.numpy(), oh.numpy()

        def e2e_step(i):
            pg.random_polyagamma(hn, zn, out=on, random_state=(777, i), disable_checks=True, first_index=first_index)
            return _lib.lib().pgm_cuda_last_host_fill_ms()

        for i in range(max(args.warmup, 1)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        dev_ms = 0.0
        for i in range(args.steps):
            dev_ms += e2e_step(100 + i)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        # the call blocks until `out` is complete; its device-side clock (first copy queued ->
        # last copy done) and the host clock around it are both reported, the slower one counts
        e2e_ms = max_over_ranks(max(dev_ms, wall_ms))
        e2e = {"value": world * ne * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 16 * ne,
               "d2h_bytes_per_step": 8 * ne, "elems_per_gpu": ne, "ms_per_step": e2e_ms / args.steps,
               "device_ms_per_step": dev_ms / args.steps, "host_ms_per_step": wall_ms / args.steps,
               "api": "polyagamma_b200.random_polyagamma(h_host, z_host, out=out_host) -> "
                      "pgm_cuda_random_polyagamma_fill2_host (pinned buffers)",
               "result_finite": bool(torch.isfinite(oh[:: max(1, ne // 1_000_000)]).all())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "elems_per_gpu": n, "global_elems": world * n,
                       "parallelism": f"index-range sharding x{world}, no collective",
                       "l2": "inputs are 16 GB per step per GPU (>> 126 MB L2): no flush needed",
                       "method_mix_rank0": mix, "seed": 12345},
            "clocks": clock_info, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "result_finite": finite,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""gpurun_out/r2_costs_<case>.csv (tools/ncu_costs.sh) -> profiles/r2_kernel_costs.json (what bench.py's roofline
reads: executed FP64 flops and DRAM bytes per UNIT OF THE CALL, per kernel kind) and profiles/r2_kernel_table.md
(every kernel of every config: time share, lanes, FP64 pipe, issue slots, flops and DRAM bytes per unit)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"cfg1": ("cfg1", 1_000_000), "cfg2": ("cfg2-scalar-h", 33554432), "cfg2_ones": ("cfg2", 33554432),
         "cfg3_devroye": ("cfg3-devroye", 33554432), "cfg3_alternate": ("cfg3-alternate", 33554432),
         "cfg3_saddle": ("cfg3-saddle", 33554432), "cfg3_gamma": ("cfg3-gamma", 33554432 // 16), "cfg4": ("cfg4", 33554432),
         "pdf": ("cfg5a-pdf", 5792 ** 2), "logpdf": ("cfg5a-logpdf", 5792 ** 2), "cdf": ("cfg5a-cdf", 5792 ** 2),
         "logcdf": ("cfg5a-logcdf", 5792 ** 2)}


def kind_of(name):
    if "tile_kernel" in name:
        return "tile"
    if "small_kernel" in name:
        return "small"
    if "density_kernel" in name:
        return "density"
    if "dense_kernel<0>" in name:
        return "gamma"
    if "dense_kernel" in name:
        return "normal"
    m = re.search(r"(setup|sample)_kernel<(\w+?)(?:T)?<?(\d)?", name)
    which = "devroye" if "Devroye" in name else "alternate_left" if "AlternateLeft" in name else "alternate" if "Alternate" in name \
        else "saddle" if "SaddleT<2>" in name or "SaddleFull" in name else "saddle_left" if "Saddle" in name else None
    if which is None:
        return None
    return ("setup_" if "setup_kernel" in name else "") + which


costs, table = {}, []
for case, (config, n) in CASES.items():
    path = os.path.join(ROOT, "gpurun_out", f"r2_costs_{case}.csv")
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        kind = kind_of(r[col["Kernel Name"]])
        if kind is None:
            continue
        val = float(r[col["Metric Value"]].replace(",", ""))
        per[(kind, r[col["ID"]])][r[col["Metric Name"]]] = val
        per[(kind, r[col["ID"]])]["name"] = r[col["Kernel Name"]]
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    for (kind, _), m in per.items():
        a = agg[kind]
        t = m["gpu__time_duration.sum"]
        a["ns"] += t
        a["launches"] += 1
        a["flops"] += 2 * m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] + \
            m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
        a["dram"] += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        a["inst"] += m["smsp__inst_executed.sum"]
        for key, src in (("pipe", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                         ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"),
                         ("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active")):
            a[key] += m[src] * t  # time-weighted
        a["regs"] = max(a["regs"], m["launch__registers_per_thread"])
    total_ns = sum(a["ns"] for a in agg.values())
    costs[config] = {}
    for kind, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        if a["ns"] < 0.002 * total_ns and a["flops"] == 0:
            continue
        costs[config][kind] = {"flops_per_unit": a["flops"] / n, "dram_bytes_per_unit": a["dram"] / n,
                               "share_of_units": 1.0, "share_of_time": a["ns"] / total_ns}
        table.append((config, kind, n, a["launches"], a["ns"] / 1e3, 100 * a["ns"] / total_ns, a["lanes"] / a["ns"], a["pipe"] / a["ns"],
                      a["issue"] / a["ns"], a["flops"] / n, a["dram"] / n, a["flops"] / a["ns"] / 1e3, a["dram"] / a["ns"], int(a["regs"])))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(costs, open(os.path.join(ROOT, "profiles", "r2_kernel_costs.json"), "w"), indent=1)
with open(os.path.join(ROOT, "profiles", "r2_kernel_table.md"), "w") as f:
    f.write("| config | kernel | units | launches | µs | share | lanes | FP64 pipe % | issue % | flops/unit | DRAM B/unit | TFLOP/s | GB/s | regs |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for t in table:
        f.write("| %s | %s | %d | %d | %.0f | %.1f %% | %.1f | %.1f | %.1f | %.0f | %.1f | %.2f | %.0f | %d |\n" % t)
print(open(os.path.join(ROOT, "profiles", "r2_kernel_table.md")).read())

"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`) per kernel.

    python tools/ncu_launch_list.py gpurun_out/launches.csv "<command that was profiled>" > profiles/<name>.txt
"""
import collections
import csv
import sys


def main(path, cmd):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")[:100]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ns = v * (1e3 if u.startswith("us") else 1e6 if u.startswith("ms") else 1e9 if u in ("s", "second") else 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    ours = {k: a for k, a in agg.items() if "pgm::" in k and "dfma_peak" not in k}
    tot = sum(a[1] for a in ours.values())
    print(f"# ncu launch list of `{cmd}`")
    print("# metric gpu__time_duration.sum, --clock-control none; per-launch times are cold-cache and")
    print("# serialised by the profiler: compare SHARES, not absolutes")
    print(f"# {len(rows)} launches, {sum(a[0] for a in ours.values())} of them kernels of libpgm_cuda.so")
    print("kernel | launches | total ms | share of libpgm_cuda kernels")
    for k, a in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"{k} | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}%")
    other = sum(a[1] for k, a in agg.items() if k not in ours)
    print(f"(other kernels in the process -- torch input generation, DFMA peak probe: {other / 1e6:.3f} ms)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")

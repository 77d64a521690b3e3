"""Executed-instruction mix of a profiled kernel (warp instructions by opcode), from the SASS view of an
.ncu-rep captured with --set full --import-source on.

    python tools/ncu_opmix.py gpurun_out/x.ncu-rep <kernel-substring> [--top 30]
"""
import argparse
import collections
import csv
import io
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("kernel")
ap.add_argument("--top", type=int, default=30)
args = ap.parse_args()
raw = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
blk = next(b for b in blocks if args.kernel in b["name"])
col = {h: i for i, h in enumerate(blk["hdr"])}
mix = collections.Counter()
for r in blk["rows"]:
    src = r[col["Source"]].strip()
    parts = src.split()
    if parts and parts[0].startswith("@"):
        parts = parts[1:]
    op = parts[0].split(".")[0] if parts else "?"
    if op in ("IMAD",) and ".WIDE" in parts[0]:
        op = "IMAD.WIDE"
    if op == "IMAD" and ".MOV" in parts[0]:
        op = "IMAD.MOV"
    mix[op] += int(r[col["Instructions Executed"]] or 0)
tot = sum(mix.values())
print(f"# {blk['name']}: {tot} warp instructions")
fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"# FP64-pipe instructions {100 * fp64 / tot:.1f} %")
for k, v in mix.most_common(args.top):
    print(f"  {k:12s} {100 * v / tot:6.2f} %")

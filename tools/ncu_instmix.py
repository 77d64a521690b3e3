"""Instruction mix (by opcode) and stall-sample share of one kernel in an .ncu-rep captured with
--import-source on.   python tools/ncu_instmix.py <rep> <kernel-regex>"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k",
                      "regex:" + rx], capture_output=True, text=True).stdout
ops = collections.Counter()
samples = collections.Counter()
tot = stot = 0
ci = None
for r in csv.reader(io.StringIO(raw)):
    if len(r) >= 2 and r[0] == "Kernel Name":
        print("kernel:", r[1])
        continue
    if len(r) > 5 and r[0] == "Address":
        ci = {h: i for i, h in enumerate(r)}
        continue
    if ci is None or len(r) < 10:
        continue
    toks = r[ci["Source"]].strip().split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    try:
        n = int(r[ci["Instructions Executed"]])
        s = int(r[ci["# Samples"]])
    except ValueError:
        continue
    ops[op] += n
    samples[op] += s
    tot += n
    stot += s
print("warp instructions executed:", tot, " stall samples:", stot)
for op, n in ops.most_common(26):
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%   samples {100 * samples[op] / max(stot, 1):5.1f}%")

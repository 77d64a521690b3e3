"""Where does a profiled kernel spend its samples?  Joins the per-instruction warp-state samples of an
.ncu-rep (`--page source`, SASS view; captured with --set full --import-source on) with the line table of
the library's cubin (`nvdisasm -g`), and prints the hottest source lines and the stall mix of each.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep <kernel-substring> [--top 40] [--launch 0]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("kernel")
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--lib", default=os.path.join(ROOT, "polyagamma_b200", "libpgm_cuda.so"))
ap.add_argument("--by", default="line", choices=["line", "file", "window", "func"])
args = ap.parse_args()

raw = subprocess.run(["ncu", "-i", args.rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
# the page is one block per launch: "Kernel Name",<name> / header / rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
def norm(name):
    name = name.replace("(bool)1", "true").replace("(bool)0", "false")
    name = re.sub(r"\((?:int|unsigned int)\)(\d+)", r"\1", name)
    return re.sub(r"\s+", "", name)


blk = next(b for b in blocks if args.kernel in b["name"])
col = {h: i for i, h in enumerate(blk["hdr"])}
base = int(blk["rows"][0][col["Address"]], 16)

# line table of the kernel from the cubin
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", args.lib], cwd=tmp, capture_output=True)
lines = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    mang = None
    where = ("?", 0)
    active = False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            mang = m.group(1)
            dem = subprocess.run(["c++filt", mang], capture_output=True, text=True).stdout.strip()
            active = norm(dem) == norm(blk["name"])
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            where = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            lines.setdefault(int(m.group(1), 16), where)

# enclosing function of a source line (rough: a header line followed by a line that is just "{")
_func_cache = {}


def func_of(fname, line):
    if fname not in _func_cache:
        starts = []
        path = os.path.join(ROOT, "polyagamma_b200", "csrc", fname)
        if os.path.exists(path):
            src = open(path).read().splitlines()
            for i, ln in enumerate(src):
                if ln.strip() == "{" and i > 0:
                    for back in range(1, 4):
                        if i - back < 0:
                            break
                        m = re.search(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", src[i - back])
                        if m and not src[i - back].rstrip().endswith(";"):
                            if m.group(1) not in ("if", "for", "while", "switch", "__launch_bounds__"):
                                starts.append((i - back + 1, m.group(1)))
                            break
        _func_cache[fname] = starts
    name = "?"
    for st, nm in _func_cache[fname]:
        if st <= line:
            name = nm
        else:
            break
    return f"{fname}:{name}"


STALLS = [h for h in blk["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
total = collections.Counter()
for r in blk["rows"]:
    off = int(r[col["Address"]], 16) - base
    where = lines.get(off, ("?", 0))
    if args.by == "file":
        key = where[0]
    elif args.by == "func":
        key = func_of(*where)
    elif args.by == "window":
        key = "0x%05x" % (off & ~0x3ff)
    else:
        key = "%s:%d" % where
    a = agg[key]
    a["samples"] += int(r[col["# Samples"]] or 0)
    a["inst"] += int(r[col["Instructions Executed"]] or 0)
    a["thr"] += int(r[col["Thread Instructions Executed"]] or 0)
    a["n"] += 1
    for s in STALLS:
        a[s] += int(r[col[s]] or 0)
for a in agg.values():
    total.update(a)
print(f"# {blk['name']}")
print(f"# samples {total['samples']}  warp-instructions {total['inst']}  lanes/inst {total['thr'] / max(total['inst'], 1):.1f}  sass-instructions {total['n']} ({total['n'] * 16 / 1024:.0f} KB)")
print("# stall mix: " + "  ".join(f"{s[6:]} {100 * total[s] / max(total['samples'], 1):.1f}%" for s in sorted(STALLS, key=lambda s: -total[s])[:9]))
print(f"{'where':38s} {'samples%':>8s} {'inst%':>6s} {'lanes':>5s}  top stalls")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[: args.top]:
    top = sorted(STALLS, key=lambda s: -a[s])[:3]
    print(f"{key:38s} {100 * a['samples'] / total['samples']:8.2f} {100 * a['inst'] / max(total['inst'], 1):6.2f} "
          f"{a['thr'] / max(a['inst'], 1):5.1f}  " + "  ".join(f"{s[6:]} {100 * a[s] / max(a['samples'], 1):.0f}%" for s in top))

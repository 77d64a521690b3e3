"""Static code-size breakdown of one kernel by source line (nvdisasm -g line table).

    python tools/sass_lines.py <kernel-substring> [--top 40] [--lib path] [--by line|file]
"""
import argparse
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("kernel")
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--lib", default=os.path.join(ROOT, "polyagamma_b200", "libpgm_cuda.so"))
ap.add_argument("--by", default="line", choices=["line", "file"])
args = ap.parse_args()
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", args.lib], cwd=tmp, capture_output=True)
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cnt, name, where = None, None, "?"
    out = {}
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cnt = out.setdefault(dem, collections.Counter()) if args.kernel in dem else None
            continue
        if cnt is None:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            where = os.path.basename(m.group(1)) + (":" + m.group(2) if args.by == "line" else "")
            continue
        if re.match(r"\s*/\*[0-9a-f]+\*/\s+\S", ln):
            cnt[where] += 1
    for dem, c in out.items():
        tot = sum(c.values())
        print(f"# {dem}: {tot} instructions, {tot * 16 / 1024:.1f} KB")
        for k, v in c.most_common(args.top):
            print(f"  {k:40s} {v:6d}  {v * 16 / 1024:6.1f} KB")

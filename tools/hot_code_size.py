#!/usr/bin/env python
"""Static code size per source function of one kernel, split into hot and cold bytes by the executed-instruction counts of an
ncu source page (measurement aid for instruction-cache work).

  python tools/hot_code_size.py dis.txt sass.csv [hot-threshold-fraction]

dis.txt / sass.csv as produced by tools/profile_by_function.sh (kept in /tmp)."""
import csv, re, sys, os
from collections import defaultdict
dis, sass = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
loc = []
cur = ("?", 0); group_open = False; found = False
for l in open(dis, errors="replace"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if not group_open:
            group_open, found = True, False
        if not found and "/rays_b200/" in m.group(1):
            cur, found = (m.group(1).split("/")[-1], int(m.group(2))), True
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        loc.append(cur); group_open = False
rows = list(csv.reader(open(sass)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
for i, r in enumerate(data):
    if r and r[0] == "Kernel Name":
        data = data[:i]; break
src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rays_b200", "csrc")
fn_of = {}
for f in os.listdir(src_dir):
    if not f.endswith((".cuh", ".cu")): continue
    name = "<top>"
    for n, l in enumerate(open(os.path.join(src_dir, f), errors="replace"), 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:RD_INLINE|RD_NOINLINE|__global__|__device__)[^(]*?(\w+)\s*\(", l)
        if m and not l.startswith(" "): name = m.group(1)
        m2 = re.match(r"^__global__.*\s(\w+)\(const TraceArgs", l)
        if m2: name = m2.group(1)
        fn_of[(f, n)] = name
ex = [float(r[col["Instructions Executed"]] or 0) for r in data]
mx = sorted(ex)[int(0.98 * len(ex))]
agg = defaultdict(lambda: [0, 0, 0.0])
for (f, ln), e in zip(loc, ex):
    k = f"{f}:{fn_of.get((f, ln), f)}"
    agg[k][0] += 16
    if e >= thr * mx: agg[k][1] += 16
    agg[k][2] += e
tot = sum(a[0] for a in agg.values()); hot = sum(a[1] for a in agg.values()); te = sum(a[2] for a in agg.values())
print(f"total {tot/1024:.1f} KB, hot (>= {thr} of the 98th-percentile count) {hot/1024:.1f} KB")
print(f"{'function':46s} {'KB':>7s} {'hot KB':>7s} {'inst%':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:46]:46s} {a[0]/1024:7.2f} {a[1]/1024:7.2f} {100*a[2]/te:6.2f}")

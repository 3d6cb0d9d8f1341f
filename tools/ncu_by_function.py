#!/usr/bin/env python
"""Aggregate an ncu SASS source page per source function / file.

  cuobjdump -xelf all build/obj/tu_2_2.o ; nvdisasm -gi trace_tu.sm_100a.cubin > all_dis.txt
  ncu -i report.ncu-rep --page source --csv --print-source sass > sass.csv
  python tools/ncu_by_function.py all_dis.txt sass.csv '<mangled kernel name>'

nvdisasm -gi annotates every instruction with the (inlined) source location; the n-th instruction of the kernel in the ncu
CSV is the n-th instruction of the kernel's section.  Prints, per innermost inlined function (and per file:line range for the
kernel body), warp instructions executed, average active lanes and the stall samples by reason."""
import csv
import re
import sys
from collections import defaultdict

dis, sass, kern = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(dis, errors="replace").read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
loc = []            # per instruction: (function, file, line)
cur_fn, cur_file, cur_line = "<kernel>", "?", 0
group_open = False   # annotations of one instruction group come innermost-first; keep the innermost location inside this repo
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if not group_open:
            group_open, found = True, False
        if not found and "/rays_b200/" in m.group(1):
            cur_file, cur_line, found = m.group(1).split("/")[-1], int(m.group(2)), True
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        loc.append((cur_file, cur_line))
        group_open = False
rows = list(csv.reader(open(sass)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
for i, r in enumerate(data):      # a report with several captured launches repeats the header block: keep the first launch
    if r and r[0] == "Kernel Name":
        data = data[:i]
        break
assert abs(len(data) - len(loc)) <= 2, (len(data), len(loc))
# map file:line -> function name by scanning the source for function headers
import os
src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rays_b200", "csrc")
fn_of = {}
for f in os.listdir(src_dir):
    if not f.endswith((".cuh", ".cu")):
        continue
    name = "<top>"
    for n, l in enumerate(open(os.path.join(src_dir, f), errors="replace"), 1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:RD_INLINE|RD_NOINLINE|__global__|__device__)[^(]*?(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            name = m.group(1)
        m2 = re.match(r"^__global__.*\s(\w+)\(const TraceArgs", l)
        if m2:
            name = m2.group(1)
        fn_of[(f, n)] = name
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: defaultdict(float))
for (f, ln), r in zip(loc, data):
    fn = fn_of.get((f, ln), f)
    key = f"{f}:{fn}"
    a = agg[key]
    a["inst"] += float(r[col["Instructions Executed"]] or 0)
    a["thr"] += float(r[col["Thread Instructions Executed"]] or 0)
    a["samples"] += float(r[col["# Samples"]] or 0)
    for s in stall_cols:
        a[s] += float(r[col[s]] or 0)
tot_i = sum(a["inst"] for a in agg.values())
tot_s = sum(a["samples"] for a in agg.values())
print(f"{'function':46s} {'inst%':>6s} {'lanes':>6s} {'samp%':>6s}  top stalls")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    if a["inst"] == 0:
        continue
    top = sorted(((a[s], s[6:]) for s in stall_cols), reverse=True)[:4]
    print(f"{k[:46]:46s} {100*a['inst']/tot_i:6.2f} {a['thr']/a['inst']:6.1f} {100*a['samples']/max(tot_s,1):6.2f}  " +
          ", ".join(f"{n} {100*v/max(a['samples'],1):.0f}%" for v, n in top))

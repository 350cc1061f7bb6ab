#!/usr/bin/env python
"""Static footprint of the HOT part of a kernel by source line: SASS instructions executed by at least `frac` x (median executed count
of the loop body) -- what competes for the SM's 32 KB L1.5 instruction cache every timestep.
usage: tools/ncu_hot_footprint.py prof.ncu-rep [min_exec]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; line = None; src = ""
rows = []
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] in ("Function Name", "Line No"): continue
    if row[0] != "":
        line = int(row[0]); src = row[1].strip(); continue
    if row[2] in ("...", "-", ""): continue
    try: n = int(row[7])
    except Exception: continue
    rows.append((cur, line, src, row[3].strip(), n))
ex = sorted(r[4] for r in rows if r[4] > 0)
med = ex[len(ex) // 2]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else med // 20
hot = [r for r in rows if r[4] >= thr]
print(f"static {len(rows)}  executed>0 {len(ex)}  median exec {med}  threshold {thr}  hot static {len(hot)} = {len(hot) * 16 / 1024:.1f} KB")
by = collections.Counter()
for r in hot: by[(r[0], r[1], r[2][:100])] += 1
for (f, l, s), n in by.most_common(45): print(f"{n:5d}  {f}:{l}  {s}")

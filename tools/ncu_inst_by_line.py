#!/usr/bin/env python
"""Executed warp-instructions per source line (ncu --import-source on, -lineinfo), largest first.
usage: tools/ncu_inst_by_line.py gpurun_out/prof_X.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, lines = None, None, []
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if row[0] == "": continue
    d = {}
    for k, v in zip(hdr, row):
        d.setdefault(k, v)
    try:
        lines.append((cur, int(row[0]), row[1].strip(), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])))
    except Exception:
        pass
ti = sum(l[3] for l in lines)
byfile = {}
for l in lines: byfile[l[0]] = byfile.get(l[0], 0) + l[3]
print("total warp inst", ti, {k: "%.1f%%" % (100 * v / ti) for k, v in byfile.items()})
lines.sort(key=lambda l: -l[3])
cum = 0
for l in lines[:N]:
    cum += l[3]
    print(f"{l[0]:16s}:{l[1]:4d} inst {100*l[3]/ti:5.2f}% cum {100*cum/ti:5.1f}% lanes {l[4]/max(1,l[3]):5.1f} | {l[2][:110]}")

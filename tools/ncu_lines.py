#!/usr/bin/env python
"""Top source lines of the step kernel by stall samples (ncu --import-source on, -lineinfo).
usage: tools/ncu_lines.py gpurun_out/prof_X.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
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
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "-")}
        lines.append((cur, int(row[0]), row[1].strip(), int(d["Instructions Executed"]), int(d["# Samples"]), stalls, d.get("L1 Wavefronts Shared", "0"), d.get("L1 Wavefronts Shared Ideal", "0")))
    except Exception:
        pass
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print("total inst", ti, "samples", ts)
agg = {}
for l in lines:
    for k, v in l[5].items(): agg[k] = agg.get(k, 0) + v
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]})
lines.sort(key=lambda l: -l[4])
for l in lines[:N]:
    top = sorted(l[5].items(), key=lambda kv: -kv[1])[:3]
    print(f"{l[0]:16s}:{l[1]:4d} smp {100*l[4]/ts:5.2f}% inst {100*l[3]/ti:5.2f}% wf {l[6]:>9s}/{l[7]:>9s} {top} | {l[2][:90]}")

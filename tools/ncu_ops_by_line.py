#!/usr/bin/env python
"""Which source lines execute a given SASS opcode most (e.g. BSSY, BRA, LDS, MOV).
usage: tools/ncu_ops_by_line.py rep OPC [N]"""
import csv, io, re, subprocess, sys
rep, opc = sys.argv[1], sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, line, agg = None, None, None, {}
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; iE = [i for i, h in enumerate(hdr) if h == "Instructions Executed"][0]; continue
    if row[0] != "":
        line = (cur, row[0], row[1].strip()[:100]); continue
    sass = re.sub(r"^@!?U?P\d+\s+", "", row[3].strip())
    if sass.split(" ")[0].split(".")[0] == opc:
        try: agg[line] = agg.get(line, 0) + int(row[iE])
        except Exception: pass
tot = sum(agg.values())
print("total", opc, tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:N]:
    print(f"{100*v/tot:5.1f}% {k[0]}:{k[1]} | {k[2]}")

#!/usr/bin/env python
"""Static SASS statistics of one kernel of libvxgpu.so: instruction count and opcode histogram.
usage: tools/sass_stats.py <substring of the mangled kernel name> [lib]"""
import collections, re, subprocess, sys
lib = sys.argv[2] if len(sys.argv) > 2 else "evosoro_b200/lib/libvxgpu.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); continue
    if cur and sys.argv[1] in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m:
            ins = re.sub(r"^@!?U?P[0-9T]+\s+", "", m.group(1).strip())
            op = ins.split()[0]
            hist[op if "--full" in sys.argv else op.split(".")[0]] += 1
tot = sum(hist.values())
print(tot, "TOTAL")
for k, v in hist.most_common(40):
    print(f"{v:6d} {k}")

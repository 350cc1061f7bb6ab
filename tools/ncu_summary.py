#!/usr/bin/env python
"""Summarise a .ncu-rep of the step kernel: headline metrics, stall reasons, executed-opcode histogram.
usage: tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [voxel_steps_per_launch]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
units_per_launch = float(sys.argv[2]) if len(sys.argv) > 2 else 2.96e7
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
for k in keys:
    if k in M:
        print(f"{k:90s} {M[k][0]:>18s} {M[k][1]}")
for h in sorted(M):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        print(f"{h:90s} {M[h][0]:>18s}")
inst = float(M["smsp__inst_executed.sum"][0])
print(f"warp-instructions per voxel-timestep: {inst / units_per_launch:.2f}  (x32 = {32 * inst / units_per_launch:.0f} thread slots)")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
cnt = collections.Counter()
static = 0
for r in rows[2:]:
    try:
        n = int(r[iE])
    except Exception:
        continue
    static += 1
    s = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
    cnt[(s.split()[0] if s else "?").split(".")[0] if "--full" not in sys.argv else s.split()[0]] += n
tot = sum(cnt.values())
print(f"static SASS instructions: {static}")
print("opcode         executed      %   thread-slots per voxel-timestep")
for k, v in cnt.most_common(32):
    print(f"{k:12s} {v:12d} {100 * v / tot:6.2f}% {v * 32 / units_per_launch:8.1f}")

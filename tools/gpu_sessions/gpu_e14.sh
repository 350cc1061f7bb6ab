#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_nan.py ex_phase_offset_full ex_lw_hopper_full 2>&1 | grep -E "barrier" | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_e14_full.log 2>&1; tail -6 gpurun_out/pytest_e14_full.log | cut -c1-250
timeout 900 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_e14.json 2> gpurun_out/bench_e14.err; tail -3 gpurun_out/bench_e14.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_e14.json"))
print("value %.4e e2e %.4e (%.3f of value) frac %.3f ms %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["roofline"]["frac"], d["ms_per_step"]))
for c in d["configs"] or []:
    print("  %-60s value %.4e frac %.3f e2e %.4e (%.1f s) %s" % (c["config"][:60], c["value"], c["roofline_frac"], c["e2e"], c["wall_s"], c["kernel"].get("strides")))
PY

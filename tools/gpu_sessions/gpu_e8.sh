#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_e8_full.log 2>&1; tail -12 gpurun_out/pytest_e8_full.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench_e8.json 2> gpurun_out/bench_e8.err; tail -3 gpurun_out/bench_e8.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_e8.json"))
print("value %.4e e2e %.4e (%.3f of value) frac %.3f ms %.2f cpu %s" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["roofline"]["frac"], d["ms_per_step"], d["cpu_baseline"] and d["cpu_baseline"]["value"]))
for c in d["configs"] or []:
    print("  %-60s value %.4e frac %.3f e2e %.4e (%.1f s)" % (c["config"][:60], c["value"], c["roofline_frac"], c["e2e"], c["wall_s"]))
PY

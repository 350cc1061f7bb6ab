#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_population.py tests/test_evaluate_all.py -m gpu -q > gpurun_out/pytest_e16.log 2>&1; tail -4 gpurun_out/pytest_e16.log | cut -c1-250
timeout 900 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_e16.json 2> gpurun_out/bench_e16.err; tail -3 gpurun_out/bench_e16.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_e16.json"))
print("value %.4e e2e %.4e (%.3f of value) frac %.3f ms %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["roofline"]["frac"], d["ms_per_step"]))
for c in d["configs"] or []:
    print("  %-60s value %.4e frac %.3f e2e %.4e (%.1f s) %s" % (c["config"][:60], c["value"], c["roofline_frac"], c["e2e"], c["wall_s"], c["kernel"].get("strides")))
PY
timeout 600 python tools/ga_generation.py --robots 1024 --ref-sample 0 > gpurun_out/ga_e16.log 2>&1; grep -E "voxelyze:|wall_s|results" gpurun_out/ga_e16.log

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_e10_full.log 2>&1; tail -12 gpurun_out/pytest_e10_full.log | cut -c1-250
timeout 900 python bench.py --steps 5 > gpurun_out/bench_e10.json 2> gpurun_out/bench_e10.err; tail -3 gpurun_out/bench_e10.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_e10.json"))
print("value %.4e e2e %.4e (%.3f of value) frac %.3f ms %.2f cpu %s" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["roofline"]["frac"], d["ms_per_step"], d["cpu_baseline"] and d["cpu_baseline"]["value"]))
for c in d["configs"] or []:
    print("  %-60s value %.4e frac %.3f e2e %.4e (%.1f s)" % (c["config"][:60], c["value"], c["roofline_frac"], c["e2e"], c["wall_s"]))
PY
for ARGS in "--lattice 6 --robots 1024 --selfcol" "--lattice 6 --robots 4096 --selfcol" "--lattice 8 --robots 2048 --workload water --selfcol"; do timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-configs $ARGS > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('$ARGS value %.4e frac %.3f e2e %.4e'%(d['value'], d['roofline']['frac'], d['e2e']['value']))" || tail -3 gpurun_out/b.err; done
timeout 900 python tools/ga_generation.py --robots 1024 --ref-sample 32 --out gpurun_out/ga_generation.json > gpurun_out/ga_generation.log 2>&1; tail -30 gpurun_out/ga_generation.log

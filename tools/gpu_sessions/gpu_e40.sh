#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "swimmer or fluid or lw_" > gpurun_out/pytest_e40.log 2>&1; tail -5 gpurun_out/pytest_e40.log | cut -c1-300
run() {
  timeout 300 python bench.py --no-cpu-baseline --no-configs --steps 5 $1 > gpurun_out/bench_e40.json 2> gpurun_out/bench_e40.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e40.json"))
    print("$1 value %.4e e2e %.4e frac %.3f ms %.2f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/bench_e40.err").read()[-600:])
PY
}
run "--lattice 8 --robots 2048 --workload water"
run "--lattice 8 --robots 2048 --workload water --selfcol"

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fluid or lw_ or gather_form" > gpurun_out/pytest_e24.log 2>&1; tail -4 gpurun_out/pytest_e24.log | cut -c1-300
run() {
  timeout 300 python bench.py --no-cpu-baseline --no-configs --steps 3 $1 > gpurun_out/bench_e24.json 2> gpurun_out/bench_e24.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e24.json"))
    print("$1 value %.4e e2e %.4e frac %.3f ms %.2f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/bench_e24.err").read()[-600:])
PY
}
for SHAPE in "--lattice 8 --robots 2048 --workload water" "--lattice 8 --robots 2048 --workload water --selfcol" "--lattice 6 --robots 1024 --selfcol" "--lattice 6 --robots 1024 --selfcol --sparse 0.55"; do
  run "$SHAPE"
done 2>&1 | tee -a gpurun_out/exp_e24.txt

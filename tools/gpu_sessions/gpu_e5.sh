#!/bin/bash
# A/B of the production kernel's forms: VXG_KERNEL_FORM = flow (default) | barrier | push
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_form or divergence or full_cli_run" > gpurun_out/pytest_e5.log 2>&1; tail -12 gpurun_out/pytest_e5.log | cut -c1-250
for SHAPE in "--lattice 10 --robots 1024" "--lattice 8 --robots 2048" "--lattice 6 --robots 4096"; do
  for FORM in $FORMS; do
    VXG_KERNEL_FORM=$FORM timeout 300 python bench.py --no-cpu-baseline --steps 5 $SHAPE > gpurun_out/bench_e5.json 2> gpurun_out/bench_e5.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e5.json"))
    print("$FORM $SHAPE value %.4e e2e %.4e frac %.3f ms %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$FORM $SHAPE FAILED", e); print(open("gpurun_out/bench_e5.err").read()[-600:])
PY
  done
done 2>&1 | tee gpurun_out/exp_e5.txt

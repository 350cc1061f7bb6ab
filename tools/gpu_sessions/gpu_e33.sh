#!/bin/bash
# Scale tabulated ahead of the launch (vx_scale_kernel): bit identity against the push form, then the shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_population.py -m gpu -x -q -k "gather_form or short_horizon or population or dev_ or example" > gpurun_out/pytest_e33.log 2>&1; tail -4 gpurun_out/pytest_e33.log | cut -c1-300
run() {
  timeout 300 python bench.py --no-cpu-baseline --no-configs --steps 5 $1 > gpurun_out/bench_e33.json 2> gpurun_out/bench_e33.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e33.json"))
    print("$1 value %.4e e2e %.4e frac %.3f ms %.2f launches %d" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["gpu_launches"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/bench_e33.err").read()[-600:])
PY
}
for SHAPE in "--lattice 10 --robots 1024" "--lattice 6 --robots 1024" "--lattice 8 --robots 2048" "--lattice 6 --robots 1024 --selfcol --sparse 0.55" "--lattice 6 --robots 1024 --selfcol" "--lattice 8 --robots 2048 --workload water" "--lattice 10 --robots 4096"; do
  run "$SHAPE"
done 2>&1 | tee -a gpurun_out/exp_e33.txt

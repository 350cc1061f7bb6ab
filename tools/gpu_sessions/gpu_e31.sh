#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --no-cpu-baseline --no-configs --steps 5 $1 > gpurun_out/bench_e31.json 2> gpurun_out/bench_e31.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e31.json"))
    print("$1 value %.4e e2e %.4e (%.3f of value) frac %.3f ms %.2f launches %d" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["roofline"]["frac"], d["ms_per_step"], d["gpu_launches"]))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/bench_e31.err").read()[-600:])
PY
}
for SHAPE in "--lattice 10 --robots 1024" "--lattice 6 --robots 1024" "--lattice 8 --robots 2048" "--lattice 6 --robots 1024 --selfcol --sparse 0.55" "--lattice 6 --robots 1024 --selfcol" "--lattice 8 --robots 2048 --workload water"; do
  run "$SHAPE"
done 2>&1 | tee -a gpurun_out/exp_e31.txt
timeout 900 python tools/ga_generation.py --robots 1024 --resident 7 --ref-sample 0 > gpurun_out/ga_e31.log 2>&1; grep -E "voxelyze:" gpurun_out/ga_e31.log | cut -c100-230; grep -A6 "wall_s_per_generation" gpurun_out/ga_e31.log | head -8

#!/bin/bash
mkdir -p gpurun_out
run() {
  VXG_LIB=$PWD/evosoro_b200/$1/libvxgpu.so timeout 300 python bench.py --no-cpu-baseline --no-configs --steps 5 $2 > gpurun_out/bench_e37.json 2> gpurun_out/bench_e37.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e37.json"))
    print("$1 $2 value %.4e e2e %.4e frac %.3f ms %.2f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"]))
except Exception as e:
    print("$1 $2 FAILED", e); print(open("gpurun_out/bench_e37.err").read()[-600:])
PY
}
for SHAPE in "--lattice 6 --robots 1024 --selfcol --sparse 0.55" "--lattice 6 --robots 4096 --selfcol --sparse 0.55"; do
  run lib "$SHAPE"; run lib_v2 "$SHAPE"
done 2>&1 | tee -a gpurun_out/exp_e37.txt

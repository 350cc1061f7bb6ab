#!/bin/bash
# A/B of kernel builds: tools/gpu_ab.sh "<lib dirs>" ["shape1" "shape2" ...]; first runs the form bit-identity tests on the first lib
mkdir -p gpurun_out
LIBS=$1; shift
if [ $# -eq 0 ]; then set -- "--lattice 10 --robots 1024" "--lattice 8 --robots 2048" "--lattice 6 --robots 4096"; fi
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_form or f32_short_horizon" > gpurun_out/pytest_ab.log 2>&1; tail -5 gpurun_out/pytest_ab.log | cut -c1-250
for SHAPE in "$@"; do
  for L in $LIBS; do
    VXG_LIB=$PWD/evosoro_b200/$L/libvxgpu.so timeout 300 python bench.py --no-cpu-baseline --steps 5 $SHAPE > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ab.json"))
    print("$L $SHAPE value %.4e e2e %.4e frac %.3f ms %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$L $SHAPE FAILED", e); print(open("gpurun_out/bench_ab.err").read()[-600:])
PY
  done
done 2>&1 | tee -a gpurun_out/exp_ab.txt

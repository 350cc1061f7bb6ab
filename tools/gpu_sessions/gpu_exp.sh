#!/bin/bash
# Tuning session: the GPU test-suite, then bench.py against each kernel build listed in $LIBS (evosoro_b200/<lib>/libvxgpu.so).
# usage (under gpurun): tools/gpu_exp.sh <tag> "<lib dirs>" ["<extra bench args>"]
TAG=${1:-exp}; LIBS=${2:-lib}; ARGS=${3:-}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
  cat gpurun_out/pytest_$TAG.log
fi
for L in $LIBS; do
  for SHAPE in "--lattice 10 --robots 1024" $SHAPES; do
    VXG_LIB=$PWD/evosoro_b200/$L/libvxgpu.so timeout 600 python bench.py --no-cpu-baseline --steps 5 $SHAPE $ARGS > gpurun_out/bench_${TAG}_$L.json 2> gpurun_out/bench_${TAG}_$L.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$L.json"))
    print("$L $SHAPE value %.4e e2e %.4e frac %.3f ms %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$L $SHAPE FAILED", e); print(open("gpurun_out/bench_${TAG}_$L.err").read()[-800:])
PY
  done
done 2>&1 | tee gpurun_out/exp_$TAG.txt

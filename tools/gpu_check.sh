#!/bin/bash
# GPU session used while tuning: parity tests, a short bench, and one full ncu capture of the step kernel.
# usage (under gpurun): tools/gpu_check.sh [tag]
TAG=${1:-chk}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_$TAG.log
python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --timesteps 100 --robots 296 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
cat gpurun_out/pytest_$TAG.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.4e e2e %.4e frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]))
PY
tail -2 gpurun_out/bench_$TAG.err

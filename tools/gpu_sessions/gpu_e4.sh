#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "full_cross or pipelined or population or gather_form" > gpurun_out/pytest_e4.log 2>&1; tail -15 gpurun_out/pytest_e4.log | cut -c1-250
timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_e4.json 2> gpurun_out/bench_e4.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_e4.json"))
print("10^3 value %.4e e2e %.4e frac %.3f ms %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel_g -s 2 -c 1 -f -o gpurun_out/prof_e4 \
    python bench.py --steps 3 --warmup 3 --timesteps 100 --robots 296 --no-cpu-baseline > gpurun_out/ncu_e4.log 2>&1
tail -3 gpurun_out/ncu_e4.log

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_form" 2>&1 | tail -15 > gpurun_out/pytest_e2_gather.log
cat gpurun_out/pytest_e2_gather.log
if ! grep -q "passed" gpurun_out/pytest_e2_gather.log || grep -q "failed" gpurun_out/pytest_e2_gather.log; then echo "gather kernel test failed: benches skipped"; fi
for SHAPE in "--lattice 10 --robots 1024" "--lattice 8 --robots 2048" "--lattice 6 --robots 4096"; do
  for PUSH in 0 1; do
    if [ $PUSH = 1 ]; then export VXG_PUSH_KERNEL=1; else unset VXG_PUSH_KERNEL; fi
    timeout 300 python bench.py --no-cpu-baseline --steps 5 $SHAPE > gpurun_out/bench_e2.json 2> gpurun_out/bench_e2.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e2.json"))
    print("push=$PUSH $SHAPE value %.4e e2e %.4e frac %.3f ms %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("push=$PUSH $SHAPE FAILED", e); print(open("gpurun_out/bench_e2.err").read()[-600:])
PY
  done
done 2>&1 | tee gpurun_out/exp_e2.txt
unset VXG_PUSH_KERNEL
timeout 1500 python -m pytest tests -m gpu -q -x --maxfail=6 > gpurun_out/pytest_e2_full.log 2>&1; tail -60 gpurun_out/pytest_e2_full.log > gpurun_out/pytest_e2.log
cat gpurun_out/pytest_e2.log

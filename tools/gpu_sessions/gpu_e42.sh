#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-configs > gpurun_out/bench_e42.json 2> gpurun_out/bench_e42.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_e42.json"))
print("value %.4e e2e %.4e (%.4f of value) frac %.3f launches %d" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["roofline"]["frac"], d["gpu_launches"]))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_form_is_bit or cli_drop_in" 2>&1 | tail -2

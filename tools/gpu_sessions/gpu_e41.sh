#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_r02c.json 2> gpurun_out/bench_n8_r02c.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n8_r02c.json").read().strip().splitlines()[-1]); print(d["n_gpus"], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], d.get("scaling"), d["clocks"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_n8_r02c.err").read()[-800:])
PY

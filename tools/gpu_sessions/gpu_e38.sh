#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_r02c.json 2> gpurun_out/bench_n2_r02c.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2_r02c.json 2> gpurun_out/bench_ref_n2_r02c.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_n2_r02c.json", "gpurun_out/bench_ref_n2_r02c.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get("impl", "ours"), d["n_gpus"], "%.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], d.get("scaling"))
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-800:])
PY

"""Scratch timing run used during bring-up (not the contract bench): N robots x n^3, K steps."""
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import common
from evosoro_b200 import abi, engine

def run(n, nrob, steps, prec=abi.VXG_PREC_F32, distinct=8):
    robs = [common.robot_via_reference(common.dense_vxa(n, s)) for s in range(distinct)]
    batch = [robs[i % distinct] for i in range(nrob)]
    b = engine.Batch(batch, precision=prec)
    cfg = b.kernel_config()
    b.step(50); b.sync()
    res = []
    for k in steps:
        b.step(k); b.sync()
        ms = b.last_step_ms()
        vs = nrob * batch[0].n_vox * k / (ms * 1e-3)
        res.append((k, ms, vs))
    b.close()
    return cfg, res

if __name__ == "__main__":
    for n, nrob in ((10, 1024), (6, 1024), (8, 1024)):
        cfg, res = run(n, nrob, (1, 10, 100, 1000))
        print(json.dumps({"n": n, "robots": nrob, "cfg": cfg, "res": [{"steps": k, "ms": ms, "voxel_steps_per_s": vs} for k, ms, vs in res]}))
    cfg, res = run(10, 296, (100,), prec=abi.VXG_PREC_F64)
    print(json.dumps({"n": 10, "robots": 296, "prec": "f64", "cfg": cfg, "res": res}))

#!/usr/bin/env python
"""bench.py -- voxel-timesteps/second of the Voxelyze timestep hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json `metric`): dense 10x10x10 four-material lattices, 1024
distinct robots PER GPU (seed = global robot index), gravity + floor + thermal actuation, evosoro's
damping/material palette.  One bench "step" = one pass of the hot path over the whole batch =
`--timesteps` (default 1000) simulation timesteps of every robot, issued as ONE kernel launch
(state stays in shared memory inside a launch).

Printed JSON (one line, rank 0):
  value      voxel-timesteps/s, all GPUs, state already resident in HBM, CUDA events on the batch stream,
             max over ranks.
  e2e        same metric through the C-ABI host-buffer call vxg_run_batch_host: descriptor flattening,
             H2D of the whole population, the same number of timesteps, D2H of every voxel position.
  roofline   algorithmic HBM bytes (115 + 72*bonds/voxels bytes per voxel-timestep) per
             launch / average launch duration, against MEASURED_PEAKS.json.
  cpu_baseline  the UNMODIFIED reference engine (oracle/_ref/ref_dump, built from /root/reference at
             -O3), one single-threaded process per host core -- evosoro's own execution model -- on a
             bounded sample of the same robots.
`--impl reference` prints the same line for that CPU arm alone.

The engine itself (libvxgpu.so) is torch-free; torch is used here only for the inter-rank barrier /
max-reduction and the device synchronize the bench contract asks for.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxel-timesteps/sec"
UNIT = "voxel-timesteps/s"
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
INIT_CM_TIME = 0.1   # s: actuation onset of the synthetic inputs (1561 steps), so the timed window covers thermal actuation


def algorithmic_bytes_per_voxel_step(n_vox: int, n_bond: int) -> float:
    """14 state words + 3 flag/material bytes per voxel, 9 history words per bond, read + written once."""
    return 115.0 + 72.0 * n_bond / n_vox


SELFCOL = False  # --selfcol: SelfColEnabled=1 in every .vxa
WATER = False   # --workload water: BASELINE.json config 5 (land/water fork, fluid drag) instead of the headline config


def synthetic_vxa(n: int, seed: int) -> str:
    from evosoro_b200 import vxa
    sim = vxa.SimParams(init_cm_time=INIT_CM_TIME, stop_condition_value=5.0, self_col_enabled=1 if SELFCOL else 0)
    if WATER:   # evosoro/examples/swimming_basic.py:68-70,137-138
        env = vxa.EnvParams(fluid_environment=1, aggregate_drag_coefficient=750.0)
        return vxa.vxa_text(vxa.dense_lattice(n, seed), sim=sim, env=env)
    return vxa.vxa_text(vxa.dense_lattice(n, seed), sim=sim)


def workload_name(n: int, robots: int) -> str:
    if WATER:
        return (f"land/water fork: dense{n}^3 x {robots} robots per GPU ({n ** 3} voxels / {3 * n * n * (n - 1)} bonds each), 4 materials, "
                "fluid environment (surface drag, C_d 750), thermal actuation")
    return (f"dense{n}^3 x {robots} robots per GPU ({n ** 3} voxels / {3 * n * n * (n - 1)} bonds each), 4 materials, "
            "gravity+floor+thermal actuation, SelfCol " + ("1 (contact-bond list rebuilt on the device)" if SELFCOL else "0"))


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(lattice: int, robots: int, timesteps: int):
    """dram bytes per launch of the step kernel from the committed ncu capture, if one matches this shape."""
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        with open(p) as f:
            for e in json.load(f):
                if e["lattice"] == lattice and e["robots"] == robots and e["timesteps"] == timesteps:
                    return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference engine, one process per core
# ------------------------------------------------------------------------------------------------
def run_reference_cpu(lattice: int, warm_ts: int, timed_ts: int, procs: int, tmpdir: str, seed0: int = 0):
    """Run `procs` concurrent single-threaded reference processes (evosoro's execution model,
    evosoro/tools/evaluation.py:89-90), each on its own robot: warm_ts untimed + timed_ts timed TimeStep()
    calls (StatToCalc default, exactly voxelyzeMain's loop).  Returns (voxel_steps_per_s, wall_s, kind)."""
    if os.access(REF_DUMP, os.X_OK):
        paths = []
        for i in range(procs):
            p = os.path.join(tmpdir, f"cpu_{lattice}_{seed0 + i}.vxa")
            with open(p, "w") as f:
                f.write(synthetic_vxa(lattice, seed0 + i))
            paths.append(p)
        exe = REF_DUMP + "_lw" if WATER else REF_DUMP
        ps = [subprocess.Popen([exe, "time", p, str(warm_ts), str(timed_ts), "nocol"], stdout=subprocess.PIPE, text=True) for p in paths]
        outs = [p.communicate()[0] for p in ps]
        vox_steps, worst = 0.0, 0.0
        for o in outs:
            f = o.split()
            vox_steps += float(f[0]) * float(f[1])
            worst = max(worst, float(f[2]))
        return vox_steps / worst, worst, "reference"
    # the reference binary did not travel: time the C oracle port instead (one process per core)
    import multiprocessing as mp
    with mp.Pool(procs) as pool:
        res = pool.map(_oracle_port_worker, [(lattice, seed0 + i, warm_ts, timed_ts) for i in range(procs)])
    worst = max(r[1] for r in res)
    return sum(r[0] for r in res) / worst, worst, "port"


def _oracle_port_worker(args):
    lattice, seed, warm_ts, timed_ts = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from evosoro_b200 import engine
    m = engine.Model(text=synthetic_vxa(lattice, seed))
    o = oracle_lib.OracleSim(m.robot)
    o.step(warm_ts)
    t0 = time.perf_counter()
    o.step(timed_ts)
    return m.robot.n_vox * timed_ts, time.perf_counter() - t0


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    ts = args.ref_timesteps
    with tempfile.TemporaryDirectory() as td:
        v, wall, kind = run_reference_cpu(args.lattice, args.warmup * ts, args.steps * ts, cores, td)
    sample = (f"{cores} concurrent single-threaded reference processes, one distinct dense {args.lattice}^3 robot each, "
              f"{args.warmup}x{ts} warm-up + {args.steps}x{ts} timed TimeStep() calls")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.lattice, args.robots), "lattice": args.lattice, "robots_per_gpu": args.robots,
                   "timesteps_per_step": ts, "init_cm_time_s": INIT_CM_TIME,
                   "sample": f"bounded sample of the workload: robots 0..{cores - 1} of the population, one per host core"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        out = self.p.communicate()[0]
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


def gpu_arm(args):
    import numpy as np
    import torch
    from evosoro_b200 import abi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        # NCCL prints its version banner (and any NCCL_DEBUG output) to stdout from C when the communicator is created; rank 0
        # must print exactly one JSON line there, so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            t = torch.zeros(1, device="cuda")
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        torch.cuda.set_device(local)
    if engine.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n, R, TS, K, W = args.lattice, args.robots, args.timesteps, args.steps, args.warmup
    # ---- inputs: R distinct robots per GPU, imported from .vxa text by the product importer ------
    robots = []
    for r in range(R):
        m = engine.Model(text=synthetic_vxa(n, rank * R + r), variant=engine.VARIANT_LAND_WATER if WATER else engine.VARIANT_VOXCAD)
        robots.append(m.robot)
        m.close()
    n_vox, n_bond = robots[0].n_vox, robots[0].n_bond
    batch = engine.Batch(robots, device=local)
    cfg = batch.kernel_config()

    # ---- value: state resident in HBM, K launches of TS timesteps --------------------------------
    batch.step(1); batch.sync()           # uploads the population (untimed), first step (dt = 0 pass)
    for _ in range(W):
        batch.step(TS)
    batch.sync()
    launches0 = batch.launch_count()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    batch.mark(0)
    for _ in range(K):
        batch.step(TS)
    batch.mark(1)
    dev_ms = batch.mark_elapsed_ms(0, 1)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = batch.launch_count() - launches0
    status = batch.sync()
    if (status & abi.VXG_ST_DIVERGED).any():
        raise SystemExit("bench.py: a robot diverged during the timed region")
    if dist is not None:
        t = torch.tensor([dev_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
    units_per_step = float(world) * R * n_vox * TS
    value = units_per_step * K / (dev_ms * 1e-3)

    # ---- the trivial end-of-job exchange: per-robot fitness scalars gathered on every rank (not timed) ---
    st = batch.stats()
    fit = np.array([[s.com[0], s.com[1], s.com[2]] for s in st])
    if dist is not None:
        tf = torch.from_numpy(fit).cuda()
        allf = [torch.empty_like(tf) for _ in range(world)]
        dist.all_gather(allf, tf)
        fit = torch.cat(allf).cpu().numpy()
    fitness_checksum = float(np.sum(fit))
    batch.close()

    # ---- e2e: host descriptors -> flatten -> H2D -> TS timesteps -> D2H positions, every step -------
    eb = engine.Batch(robots, device=local, upload=False)
    pos = np.zeros((R, eb.limits.max_vox, 3))
    ke = max(1, min(K, args.e2e_steps))
    for _ in range(min(W, 2)):
        eb.run_batch_host(TS, pos)
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        eb.run_batch_host(TS, pos)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = units_per_step * ke / e2e_s
    h2d = eb.h2d_bytes_per_upload()
    stride = 64
    while stride < eb.limits.max_vox:
        stride *= 2
    for lo, hi in ((128, 224), (256, 352), (512, 736)):   # stride_for(), csrc/vx_robot.cuh
        if lo < eb.limits.max_vox <= hi:
            stride = hi
    d2h = R * stride * 3 * 8   # the position planes of every blob at the device stride (the per-robot control blocks, ~0.3 KB each, not counted)
    if not np.isfinite(pos).all():
        raise SystemExit("bench.py: non-finite positions from the e2e path")
    eb.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        warm_ts, timed_ts = 200, args.cpu_timesteps
        with tempfile.TemporaryDirectory() as td:
            v, wall, kind = run_reference_cpu(n, warm_ts, timed_ts, cores, td)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{cores} concurrent single-threaded reference processes (evosoro's execution model), one distinct dense {n}^3 robot each, "
                         f"{warm_ts} warm-up + {timed_ts} timed TimeStep() calls, {wall:.1f} s"}

    if rank == 0:
        bpv = algorithmic_bytes_per_voxel_step(n_vox, n_bond)
        launch_s = dev_ms * 1e-3 / max(1, launches)
        alg_bytes_per_launch = bpv * R * n_vox * TS * (K / max(1, launches))
        peak, peak_src = measured_peak_gbs()
        achieved = alg_bytes_per_launch / launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (forces, moments, quaternions, momenta) + f64 (positions, time, scale law)", "data": "synthetic",
            "config": {"workload": workload_name(n, R),
                       "lattice": n, "robots_per_gpu": R, "timesteps_per_step": TS, "init_cm_time_s": INIT_CM_TIME,
                       "l2": f"inputs larger than L2: {R} robot blobs x {cfg['blob_bytes']} B = {R * cfg['blob_bytes'] / 1e6:.0f} MB streamed per launch (L2 126 MB)",
                       "kernel": cfg, "parallelism": f"{world} x independent shards, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(n, R, TS), "peak_source": peak_src,
                         "algorithmic_bytes_per_voxel_step": bpv, "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "launch_ms": launch_s * 1e3, "kernel": "vx_step_kernel",
                         "note": "state is shared-memory resident inside a launch, so achieved is EFFECTIVE algorithmic bandwidth; physical DRAM bytes are in `traffic`"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": ke,
                    "call": "vxg_run_batch_host (flatten + upload + step + download positions)"},
            "gpu_launches": launches, "wall_ms_per_step": wall_ms / K, "clocks": clocks,
            "fitness_gather": {"robots": int(fit.shape[0]), "checksum": fitness_checksum},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lattice", type=int, default=10)
    ap.add_argument("--robots", type=int, default=1024, help="robots per GPU")
    ap.add_argument("--timesteps", type=int, default=1000, help="simulation timesteps per bench step (one launch)")
    ap.add_argument("--ref-timesteps", type=int, default=300, help="timesteps per bench step of the CPU reference arm")
    ap.add_argument("--cpu-timesteps", type=int, default=10000, help="timed timesteps of the cpu_baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--selfcol", action="store_true", help="SelfColEnabled=1 (evosoro's GA default): the self-collision kernel family; our arm only")
    ap.add_argument("--workload", default="land", choices=["land", "water"], help="water = BASELINE config 5 (use --lattice 8 --robots 2048)")
    args = ap.parse_args()
    global WATER, SELFCOL
    WATER = args.workload == "water"
    SELFCOL = bool(args.selfcol)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())

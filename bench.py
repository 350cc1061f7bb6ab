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
  e2e        same metric through the C-ABI host-buffer call vxg_pop_run_host: descriptor flattening on the host
             cores, H2D of the whole population, --e2e-timesteps timesteps from the reset state, D2H of every voxel
             position; pipelined over chunks of robots (flatten / upload / simulate / download overlap).
  configs    the same two legs for the other BASELINE.json configurations and for evosoro's GA default
             (6^3 workspace, CPPN-like sparse morphologies, SelfColEnabled=1); N = 1 only.
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
INIT_CM_TIME = 0.02  # s: actuation onset of the synthetic inputs (timestep 313).  Both legs simulate the same window -- timesteps [0, steps x timesteps)
                     # of a fresh population, 97 % of it thermally actuated at the default 10 x 1000


def algorithmic_bytes_per_voxel_step(n_vox: int, n_bond: int) -> float:
    """14 state words + 3 flag/material bytes per voxel, 9 history words per bond, read + written once."""
    return 115.0 + 72.0 * n_bond / n_vox


SELFCOL = False  # --selfcol: SelfColEnabled=1 in every .vxa
WATER = False   # --workload water: BASELINE.json config 5 (land/water fork, fluid drag) instead of the headline config


# the other configurations bench.py reports next to the headline (BASELINE.json configs 1-2, 3, 5 and evosoro's actual GA default)
OTHER_CONFIGS = [
    ("BASELINE config 2: 6^3 x 1024, terrestrial", dict(lattice=6, robots=1024)),
    ("BASELINE config 3: 10^3 x 4096, thermal actuation + floor", dict(lattice=10, robots=4096)),
    ("BASELINE config 5: land/water 8^3 x 2048, fluid drag", dict(lattice=8, robots=2048, water=True)),
    ("evosoro GA default: 6^3 workspace, CPPN-like morphologies >= 30 % full (examples/basic.py:114-127), SelfColEnabled=1", dict(lattice=6, robots=1024, selfcol=True, sparse=0.55)),
    # the same shapes with evosoro's default SelfColEnabled=1 (base.py:24): the self-collision kernel family
    ("6^3 x 1024 dense, SelfColEnabled=1", dict(lattice=6, robots=1024, selfcol=True)),
    ("10^3 x 1024, SelfColEnabled=1", dict(lattice=10, robots=1024, selfcol=True)),
    ("land/water 8^3 x 2048, fluid drag, SelfColEnabled=1", dict(lattice=8, robots=2048, water=True, selfcol=True)),
]


def synthetic_vxa(n: int, seed: int, water=None, selfcol=None, sparse: float = 0.0) -> str:
    from evosoro_b200 import vxa
    water = WATER if water is None else water
    selfcol = SELFCOL if selfcol is None else selfcol
    sim = vxa.SimParams(init_cm_time=INIT_CM_TIME, stop_condition_value=5.0, self_col_enabled=1 if selfcol else 0)
    # sparse > 0: a connected blob filling that fraction of the n^3 workspace (fill varies +-0.2 with the seed, never below 0.3)
    fill = max(0.3, min(1.0, sparse + 0.2 * (((seed * 2654435761) % 1000) / 500.0 - 1.0))) if sparse > 0 else 0.0
    st = vxa.sparse_lattice(n, seed, fill=fill) if sparse > 0 else vxa.dense_lattice(n, seed)
    if water:   # evosoro/examples/swimming_basic.py:68-70,137-138
        env = vxa.EnvParams(fluid_environment=1, aggregate_drag_coefficient=750.0)
        return vxa.vxa_text(st, sim=sim, env=env)
    return vxa.vxa_text(st, sim=sim)


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


def ncu_traffic_per_launch(lattice: int, robots_per_launch: int, timesteps: int):
    """dram bytes per launch of the step kernel from the committed ncu capture (bytes per robot and launch x robots in a launch)."""
    p = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        with open(p) as f:
            for e in json.load(f):
                if e["lattice"] == lattice and e["timesteps"] == timesteps:
                    return float(e["dram_bytes_per_robot_launch"]) * robots_per_launch
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference engine, one process per core
# ------------------------------------------------------------------------------------------------
def run_reference_cpu(lattice: int, warm_ts: int, timed_ts: int, procs: int, tmpdir: str, seed0: int = 0, exe_suffix: str = ""):
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
        exe = REF_DUMP + "_lw" if WATER else REF_DUMP + exe_suffix
        # (the reference arm runs with self-collision exactly when the robots ask for it: --selfcol)
        ps = [subprocess.Popen([exe, "time", p, str(warm_ts), str(timed_ts)] + ([] if SELFCOL else ["nocol"]), stdout=subprocess.PIPE, text=True) for p in paths]
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


def build_robots(n, count, seed0, water=False, selfcol=False, sparse=0.0):
    """`count` distinct synthetic robots imported from .vxa text by the product importer (seed = global robot index)."""
    from evosoro_b200 import engine
    robots = []
    for r in range(count):
        m = engine.Model(text=synthetic_vxa(n, seed0 + r, water=water, selfcol=selfcol, sparse=sparse),
                         variant=engine.VARIANT_LAND_WATER if water else engine.VARIANT_VOXCAD)
        robots.append(m.robot)
        m.close()
    return robots


def _resident_leg(robots, local, TS, K, W, barrier, allreduce_max, sampler, chunk_robots):
    """The population resident in HBM, as the same chunked batches the e2e call uses (one stream per chunk: the chunks' grids
    overlap, so a launch's last, partly filled wave of CTAs does not idle SMs); device time from CUDA events on those streams."""
    import numpy as np
    from evosoro_b200 import abi, engine
    rp = engine.Population(robots, device=local, chunk_robots=chunk_robots)
    rp.step(1); rp.sync()                 # first step (dt = 0 pass)
    for _ in range(W):                    # warm-up launches: actuated steps of the same robots
        rp.step(TS)
    rp.sync()
    rp.upload(); rp.sync()                # ResetSimulation for the whole population, upload completed here (untimed): the timed
    launches0 = rp.launch_count()         # launches cover timesteps [0, K * TS) of a fresh population, like an e2e call
    barrier()
    if sampler is not None:
        sampler.start()
    t_wall0 = time.perf_counter()
    rp.mark(0)
    for _ in range(K):
        rp.step(TS)
    rp.mark(1)
    dev_ms = rp.mark_elapsed_ms(0, 1)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop() if sampler is not None else None
    launches = rp.launch_count() - launches0
    status = rp.sync()
    if (status & abi.VXG_ST_DIVERGED).any():
        raise SystemExit("bench.py: a robot diverged during the timed region")
    dev_ms, wall_ms = allreduce_max([dev_ms, wall_ms])
    info = rp.batches()
    cfg = {"threads": info[0][2], "stride": info[0][1], "family_bits": info[0][3], "robots_per_batch": info[0][0], "batches": len(info),
           "strides": sorted({b[1] for b in info})}
    st = rp.stats()
    fit = np.array([[x.com[0], x.com[1], x.com[2]] for x in st])
    rp.close()
    return dev_ms, wall_ms, launches, clocks, cfg, fit


def measure(robots, local, TS, K, W, e2e_ts, e2e_calls, barrier, allreduce_max, chunk_robots, sampler=None, ragged=False):
    """One workload on this rank's GPU.  `value` leg: population resident in HBM; after W warm-up launches the population is
    reset (untimed) and K launches of TS timesteps are timed with CUDA events on the batch stream -- timesteps [0, K*TS) of a
    fresh population (actuation starts at timestep init_cm_time / dt).  `e2e` leg: the host-buffer call vxg_pop_run_host --
    descriptors flattened on the host cores, H2D of the population, e2e_ts timesteps from the reset state (by default the SAME
    window as the `value` leg), D2H of every voxel position -- pipelined over chunks of `chunk_robots` robots, wall clock
    around the call."""
    import numpy as np
    from evosoro_b200 import abi, engine
    R = len(robots)
    vox_total = float(sum(r.n_vox for r in robots))
    bond_total = float(sum(r.n_bond for r in robots))
    dev_ms, wall_ms, launches, clocks, cfg, fit = _resident_leg(robots, local, TS, K, W, barrier, allreduce_max, sampler, chunk_robots)

    pop = engine.Population(robots, device=local, chunk_robots=chunk_robots)
    pos = np.zeros((R, pop.max_vox, 3))
    pop.run_host(e2e_ts, pos)             # warm-up call (first-touch of the pinned mirrors, kernel attributes)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_calls):
        pop.run_host(e2e_ts, pos)
    barrier()
    e2e_s, = allreduce_max([time.perf_counter() - t0])
    if not np.isfinite(pos).all():
        raise SystemExit("bench.py: non-finite positions from the e2e path")
    out = {"dev_ms": dev_ms, "wall_ms": wall_ms, "launches": launches, "clocks": clocks, "cfg": cfg, "fit": fit,
           "vox_steps_per_step": vox_total * TS, "n_vox_mean": vox_total / R, "bonds_per_voxel": bond_total / vox_total,
           "e2e_s": e2e_s, "e2e_units": vox_total * e2e_ts * e2e_calls, "h2d": pop.h2d_bytes(), "d2h": pop.d2h_bytes(),
           "batches": len(pop.batches()), "e2e_launches": None}
    pop.close()
    return out


def gpu_arm(args):
    import numpy as np
    import torch
    from evosoro_b200 import engine

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

    def allreduce_max(vals):
        if dist is None:
            return list(vals)
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    n, R, TS, K, W = args.lattice, args.robots, args.timesteps, args.steps, args.warmup
    if args.e2e_timesteps <= 0:
        args.e2e_timesteps = K * TS
    robots = build_robots(n, R, rank * R, water=WATER, selfcol=SELFCOL, sparse=args.sparse)
    n_vox, n_bond = robots[0].n_vox, robots[0].n_bond
    sampler = ClockSampler(local) if rank == 0 else None
    m = measure(robots, local, TS, K, W, args.e2e_timesteps, max(1, min(K, args.e2e_steps)), barrier, allreduce_max, args.chunk_robots, sampler,
                ragged=args.sparse > 0)
    value = float(world) * m["vox_steps_per_step"] * K / (m["dev_ms"] * 1e-3)
    e2e_value = float(world) * m["e2e_units"] / m["e2e_s"]

    # ---- the trivial end-of-job exchange: per-robot fitness scalars gathered on every rank (not timed) ---
    fit = m["fit"]
    if dist is not None:
        tf = torch.from_numpy(fit).cuda()
        allf = [torch.empty_like(tf) for _ in range(world)]
        dist.all_gather(allf, tf)
        fit = torch.cat(allf).cpu().numpy()
    fitness_checksum = float(np.sum(fit))

    # ---- the other BASELINE.json configurations and evosoro's GA default, same legs, fewer repetitions (N = 1 only) -------
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        peak, _ = measured_peak_gbs()
        configs = []
        for name, kw in OTHER_CONFIGS:
            t0 = time.perf_counter()
            rb = build_robots(kw["lattice"], kw["robots"], 0, water=kw.get("water", False), selfcol=kw.get("selfcol", False), sparse=kw.get("sparse", 0.0))
            c = measure(rb, local, TS, 3, 3, 3 * TS, 1, barrier, allreduce_max, args.chunk_robots, ragged=kw.get("sparse", 0.0) > 0)
            v = c["vox_steps_per_step"] * 3 / (c["dev_ms"] * 1e-3)
            bpv = 115.0 + 72.0 * c["bonds_per_voxel"]
            configs.append({"config": name, **{k: kw[k] for k in kw}, "voxels_per_robot": c["n_vox_mean"], "value": v, "unit": UNIT,
                            "ms_per_step": c["dev_ms"] / 3, "roofline_frac": v * bpv / 1e9 / peak, "algorithmic_bytes_per_voxel_step": bpv,
                            "e2e": c["e2e_units"] / c["e2e_s"], "kernel": c["cfg"], "batches": c["batches"], "wall_s": time.perf_counter() - t0})
            del rb

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
        if kind == "reference" and not WATER:
            # two more data points BASELINE.md sections 3.1 / 3.3 ask for: ONE reference process (the engine is single-threaded: this
            # is what one evaluation costs), and the same with the library compiled as the reference ships it (no optimisation flags)
            with tempfile.TemporaryDirectory() as td:
                v1, w1, _ = run_reference_cpu(n, 100, max(200, timed_ts // 10), 1, td)
                cpu["single_process"] = {"value": v1, "cores": 1, "build": "-O3", "wall_s": round(w1, 2)}
                if os.access(REF_DUMP + "_O0", os.X_OK):
                    v0, w0, _ = run_reference_cpu(n, 50, max(100, timed_ts // 30), 1, td, exe_suffix="_O0")
                    cpu["as_shipped_O0"] = {"value": v0, "cores": 1, "build": "-O0 (Voxelyze/Makefile:92-95 leaves CXXFLAGS unset)", "wall_s": round(w0, 2)}

    if rank == 0:
        launches, dev_ms, cfg = m["launches"], m["dev_ms"], m["cfg"]
        bpv = algorithmic_bytes_per_voxel_step(n_vox, n_bond) if args.sparse <= 0 else 115.0 + 72.0 * m["bonds_per_voxel"]
        launch_s = dev_ms * 1e-3 / max(1, launches)
        alg_bytes_per_launch = bpv * m["vox_steps_per_step"] * (K / max(1, launches))
        peak, peak_src = measured_peak_gbs()
        achieved = alg_bytes_per_launch / launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (forces, moments, quaternions, momenta) + f64 (positions, time, scale law)", "data": "synthetic",
            "config": {"workload": workload_name(n, R),
                       "lattice": n, "robots_per_gpu": R, "timesteps_per_step": TS, "init_cm_time_s": INIT_CM_TIME,
                       "l2": f"inputs larger than L2: {R} robot blobs of {216 * cfg['stride']} B = {R * 216 * cfg['stride'] / 1e6:.0f} MB streamed through per launch wave (L2 126 MB); state on chip inside a launch",
                       "kernel": cfg, "parallelism": f"{world} x independent shards, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_per_launch(n, cfg["robots_per_batch"], TS) if not (SELFCOL or WATER or args.sparse > 0) else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_voxel_step": bpv, "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                         "launch_ms": launch_s * 1e3, "kernel": "vx_step_kernel_g" if not (SELFCOL or WATER) and cfg["stride"] <= 1024 else "vx_step_kernel",
                         "launches_overlap": f"{cfg['batches']} batches on their own streams: achieved = algorithmic bytes of all launches / device time of the region",
                         "note": "state is on-chip (shared + tensor memory) inside a launch, so achieved is EFFECTIVE algorithmic bandwidth; physical DRAM bytes are in `traffic`"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"], "steps": max(1, min(K, args.e2e_steps)),
                    "timesteps_per_step": args.e2e_timesteps,
                    "call": f"vxg_pop_run_host (flatten + upload + {args.e2e_timesteps} timesteps from the reset state + download positions), "
                            f"pipelined over {m['batches']} chunks of <= {args.chunk_robots} robots",
                    "regime": f"timesteps 0..{args.e2e_timesteps - 1} of a fresh population; actuation starts at timestep {int(INIT_CM_TIME / 6.405863e-05) + 1}; "
                              f"`value` is timed on timesteps 0..{K * TS - 1} of the same population"},
            "gpu_launches": launches, "wall_ms_per_step": m["wall_ms"] / K, "clocks": m["clocks"],
            "fitness_gather": {"robots": int(fit.shape[0]), "checksum": fitness_checksum},
            "configs": configs,
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
    ap.add_argument("--e2e-steps", type=int, default=2, help="timed e2e calls")
    ap.add_argument("--e2e-timesteps", type=int, default=0, help="timesteps per e2e call (from the reset state); 0 = steps x timesteps, the window `value` is timed on")
    ap.add_argument("--chunk-robots", type=int, default=128, help="pipelining granularity of the e2e call")
    ap.add_argument("--sparse", type=float, default=0.0, help="> 0: CPPN-like connected blobs filling about this fraction of the workspace")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (other BASELINE configurations)")
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

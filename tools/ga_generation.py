#!/usr/bin/env python
"""One GA generation through the batched front end: N .vxa files in -> N result XMLs out, whole
lifetimes, next to evosoro's own path (one `voxelyze -f` process per individual, all cores busy; evaluation.py:89-158) on a
bounded sample of the same files on the same box.

    python tools/ga_generation.py [--robots 1024] [--ref-sample 32] [--devices 0] [--sim-time 5] [--out profiles/x.json]

The population is what evosoro's stock experiment evaluates (examples/basic.py): 6^3 workspace, CPPN-like connected morphologies
at least 30 % full, 4 materials, SelfColEnabled=1, 5 s lifetime after 1 s of settling."""
import argparse, json, os, subprocess, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evosoro_b200 import evaluation, vxa  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "voxelyze_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", type=int, default=1024)
    ap.add_argument("--ref-sample", type=int, default=32)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--sim-time", type=float, default=5.0)
    ap.add_argument("--init-time", type=float, default=1.0)
    ap.add_argument("--resident", type=int, default=3, help="generations through one resident engine process (voxelyze --serve); 0 = skip")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    devices = [int(x) for x in a.devices.split(",")]
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(td + "/fit"); os.makedirs(td + "/ref")
        t0 = time.perf_counter()
        paths, vox = [], 0
        for i in range(a.robots):
            fill = 0.3 + 0.5 * ((i * 2654435761) % 1000) / 1000.0
            st = vxa.sparse_lattice(6, i, fill=fill)
            vox += int((st > 0).sum())
            sim = vxa.SimParams(init_cm_time=a.init_time, stop_condition_value=a.sim_time, self_col_enabled=1, fitness_file=f"{td}/fit/{i:05d}.xml")
            p = f"{td}/{i:05d}.vxa"
            with open(p, "w") as f:
                f.write(vxa.vxa_text(st, sim=sim))
            paths.append(p)
        t_write = time.perf_counter() - t0
        plan = subprocess.run([evaluation.VOXELYZE, "--plan"] + sum((["-f", p] for p in paths), []), capture_output=True, text=True).stdout
        work = sum(int(l.split()[1]) * int(l.split()[3]) for l in plan.strip().splitlines())
        os.environ["VXG_TIMING"] = "1"    # the CLI prints its phase times on stderr
        t0 = time.perf_counter()
        codes = evaluation.run_batch(paths, devices=devices)
        t_ours = time.perf_counter() - t0
        n_out = len(os.listdir(td + "/fit"))
        res = {"robots": a.robots, "voxels_mean": vox / a.robots, "voxel_timesteps": work, "devices": devices, "write_vxa_s": t_write,
               "ours": {"wall_s": t_ours, "results": n_out, "exit": codes, "voxel_timesteps_per_s": work / t_ours,
                        "call": "evaluation.run_batch -> voxelyze --devices ... -f x N (import on the host cores, ragged populations on the GPUs, N result files)"}}
        if a.resident > 0:   # the same generation again and again through ONE engine process: what a GA run pays from generation 2 on
            walls = []
            for g in range(a.resident):
                for f in os.listdir(td + "/fit"):
                    os.remove(td + "/fit/" + f)
                t0 = time.perf_counter()
                codes = evaluation.run_batch(paths, devices=devices, resident=True)
                walls.append(time.perf_counter() - t0)
                assert len(os.listdir(td + "/fit")) == n_out, (len(os.listdir(td + "/fit")), n_out)
            res["ours_resident"] = {"wall_s_per_generation": walls, "exit": codes, "voxel_timesteps_per_s": work / min(walls),
                                    "call": "evaluation.run_batch(resident=True) -> one `voxelyze --serve` process for the whole run; the first "
                                            "generation includes its start (CUDA context)"}
        if os.access(REF, os.X_OK) and a.ref_sample > 0:
            m = min(a.ref_sample, a.robots)
            cores = os.cpu_count() or 1
            sample = []
            for i in range(m):
                text = open(paths[i]).read().replace(f"{td}/fit/{i:05d}.xml", f"{td}/ref/{i:05d}.xml")
                p = f"{td}/ref_{i:05d}.vxa"
                open(p, "w").write(text); sample.append(p)
            w = sum(int(l.split()[1]) * int(l.split()[3]) for l in plan.strip().splitlines()[:m])
            t0 = time.perf_counter()
            running, todo = [], list(sample)
            while todo or running:       # evosoro launches everything at once and lets the OS schedule; here: one process per core
                while todo and len(running) < cores:
                    running.append(subprocess.Popen([REF, "-f", todo.pop(0)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
                running = [p for p in running if p.poll() is None]
                time.sleep(0.01)
            t_ref = time.perf_counter() - t0
            res["reference"] = {"sample_robots": m, "cores": cores, "wall_s": t_ref, "results": len(os.listdir(td + "/ref")),
                                "voxel_timesteps_per_s": w / t_ref, "projected_wall_s_for_all": t_ref * work / max(1, w),
                                "call": "oracle/_ref/voxelyze_ref -f <file>, one process per individual, one per core at a time (evaluation.py:89-90)"}
            res["speedup_projected"] = res["reference"]["projected_wall_s_for_all"] / t_ours
            if "ours_resident" in res:
                res["speedup_projected_resident"] = res["reference"]["projected_wall_s_for_all"] / min(res["ours_resident"]["wall_s_per_generation"])
    print(json.dumps(res, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

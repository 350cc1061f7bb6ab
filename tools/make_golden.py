#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference engine (oracle/_ref/ref_dump, built from
/root/reference by oracle/Makefile).  Run in the build container (needs oracle/_ref); the fixtures are
committed so that the CPU and GPU test-suites can pin the oracle / the kernels without the reference.

    python tools/make_golden.py

Each fixture holds the .vxa text, the reference's imported arrays (dump header) and sampled frames of
per-voxel state + per-bond outputs -- all fp64, exactly as the reference's public getters return them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refdump  # noqa: E402
from evosoro_b200 import vxa  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TMP = "/tmp/vx_golden"


def case(name, structure, n_steps, every, sim=None, env=None, palette=None, phase=None, lw=False, full=False, col=False, bcs=None, per_voxel=None):
    os.makedirs(TMP, exist_ok=True)
    if full:   # the stock CLI writes its result file where <FitnessFileName> says
        sim = sim or vxa.SimParams()
        sim.fitness_file = os.path.join(TMP, name + "_result.xml")
    text = vxa.vxa_text(structure, palette=palette, sim=sim, env=env, phase_offset=phase, boundary_conditions=bcs, per_voxel=per_voxel)
    case_text(name, text, n_steps, every, lw=lw, full=full, col=col, fitness_file=sim.fitness_file if full else None)


def case_text(name, text, n_steps, every, lw=False, full=False, col=False, fitness_file=None, suffix=""):
    class _S: pass
    sim = _S(); sim.fitness_file = fitness_file
    os.makedirs(TMP, exist_ok=True)
    path = os.path.join(TMP, name + ".vxa")
    with open(path, "w") as f:
        f.write(text)
    d = refdump.run_ref_dump(path, path + ".bin", n_steps, every, nocol=not col, lw=lw, suffix=suffix)
    out = dict(vxa=np.frombuffer(text.encode("latin-1"), dtype=np.uint8), lw=np.int32(lw), n_steps=np.int64(n_steps), every=np.int64(every),
               sim=d.sim, vox=d.vox, bond=d.bond, frame_hdr=d.frame_hdr, vstate=d.vstate, bstate=d.bstate)
    if full:
        f = refdump.run_ref_full(path, path + ".full.bin", nocol=not col, lw=lw, suffix=suffix)
        out.update(full_hdr=f.frame_hdr, full_vstate=f.vstate[-1], full_tail=f.tail)
        # the unmodified reference CLI (voxelyzeMain/main.cpp) on the same file: its result XML is the P5 pin
        import subprocess
        exe = os.path.join(ROOT, "oracle", "_ref", "voxelyze_ref_lw" if lw else "voxelyze_ref" + suffix)
        if os.path.exists(sim.fitness_file):
            os.remove(sim.fitness_file)
        rc = subprocess.run([exe, "-f", path]).returncode
        assert rc == 1, rc
        out.update(result_xml=np.frombuffer(open(sim.fitness_file, "rb").read(), dtype=np.uint8), fitness_file=np.frombuffer(sim.fitness_file.encode(), dtype=np.uint8))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "voxels", d.n_vox, "bonds", d.n_bond, "frames", d.vstate.shape[0], os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")


def main():
    os.makedirs(OUT, exist_ok=True)
    S = vxa.SimParams
    # settling + actuation onset inside the window (InitCmTime 0.02 s = 312 steps)
    case("dense3_s5", vxa.dense_lattice(3, 5), 1500, 50, sim=S(init_cm_time=0.02, stop_condition_value=0.2))
    case("dense4_s7", vxa.dense_lattice(4, 7), 1200, 100, sim=S(init_cm_time=0.02, stop_condition_value=0.2))
    # every step of the first 40 (dt = 0 first pass, damping history, mode flags)
    case("dense3_s2_first40", vxa.dense_lattice(3, 2), 40, 1, sim=S(init_cm_time=0.001, stop_condition_value=0.2))
    # ragged morphology (CPPN-like blob), soft-only palette -> 10x larger dt
    case("sparse5_s3", vxa.sparse_lattice(5, 3, fill=0.5), 1000, 100, sim=S(init_cm_time=0.02, stop_condition_value=0.2))
    # per-voxel phase offsets
    rng = np.random.RandomState(4)
    case("dense3_phase", vxa.dense_lattice(3, 9), 1000, 100, sim=S(init_cm_time=0.01, stop_condition_value=0.2), phase=rng.uniform(-0.5, 0.5, (3, 3, 3)))
    # whole CLI run to StopConditionMet: step count, IniCM latch, fitness block
    case("dense3_full", vxa.dense_lattice(3, 11, materials=(1, 3, 4)), 200, 100, sim=S(init_cm_time=0.05, stop_condition_value=0.25), full=True)
    # afterlife: EndOfLifetimePosteriorY is latched at StopConditionValue, the run ends AfterlifeTime later
    case("dense3_afterlife_full", vxa.dense_lattice(3, 12, materials=(1, 3, 4)), 0, 1,
         sim=S(init_cm_time=0.05, stop_condition_value=0.2, afterlife_time=0.05), full=True)
    # land/water fork, whole run, its own result tags
    case("lw_dense3_full", vxa.dense_lattice(3, 13, materials=(1, 3, 4)), 0, 1, sim=S(init_cm_time=0.05, stop_condition_value=0.25), lw=True, full=True)
    # BASELINE config 1/2 shape: dense 6^3, whole 5 s CLI run (78 054 steps): final COM / fitness only
    case("dense6_s1_full", vxa.dense_lattice(6, 1), 0, 1, full=True)
    # the headline shape: dense 10^3, whole 5 s CLI run (78 054 steps x 1000 voxels): final COM / fitness only
    case("dense10_s3_full", vxa.dense_lattice(10, 3), 0, 1, full=True)
    # materials of different density: the two voxels of a bond have different mass, so Force2 != -Force1 in the damping terms
    # (VXS_BondInternal.cpp:310-346) -- such robots run the wide kernel family
    pal = vxa.four_material_palette()
    pal[0].density, pal[1].density, pal[2].density, pal[3].density = 1e6, 3e6, 5e5, 2e6
    case("dense3_density", vxa.dense_lattice(3, 14), 1200, 100, sim=S(init_cm_time=0.02, stop_condition_value=0.2), palette=pal)
    # single voxel / no bonds (edge case: empty bond list)
    one = np.zeros((2, 2, 2), dtype=np.int32); one[0, 0, 0] = 1
    case("single_voxel", one, 300, 100, sim=S(init_cm_time=0.01, stop_condition_value=0.1))
    # land/water fork scale law, no fluid
    case("lw_dense3_s6", vxa.dense_lattice(3, 6), 1000, 100, sim=S(init_cm_time=0.01, stop_condition_value=0.2), lw=True)


def pincer(h=5):
    """Two actuated pillars on a passive base that bend towards each other: sustained self-contact between voxels
    that are far apart in the lattice graph (contact bonds, VX_Sim.cpp:2357-2413)."""
    st = np.zeros((6, 2, h + 1), dtype=np.int32)
    st[:, :, 0] = 1
    st[0, :, 1:] = 3; st[1, :, 1:] = 4; st[4, :, 1:] = 4; st[5, :, 1:] = 3
    return st


def collision_cases():
    S = vxa.SimParams
    case("col_pincer", pincer(), 800, 50, sim=S(init_cm_time=0.005, stop_condition_value=0.5, self_col_enabled=1), col=True)
    case("col_pincer_full", pincer(4), 0, 1, sim=S(init_cm_time=0.05, stop_condition_value=0.6, self_col_enabled=1), col=True, full=True)
    # SelfColEnabled=1 but nothing ever touches: list rebuilds only
    case("col_dense3", vxa.dense_lattice(3, 31), 600, 100, sim=S(init_cm_time=0.02, stop_condition_value=0.2, self_col_enabled=1), col=True)


def fluid_cases():
    """land/water fork in a fluid environment (config 5 of BASELINE.json: surface drag, gravity and floor inert)"""
    S, E = vxa.SimParams, vxa.EnvParams
    env = lambda: E(fluid_environment=1, aggregate_drag_coefficient=750.0)
    case("lw_fluid_dense3", vxa.dense_lattice(3, 21, materials=(1, 3, 4)), 1200, 100, sim=S(init_cm_time=0.01, stop_condition_value=0.2), env=env(), lw=True)
    case("lw_fluid_sparse4", vxa.sparse_lattice(4, 22, fill=0.5, materials=(1, 3, 4)), 1000, 100, sim=S(init_cm_time=0.01, stop_condition_value=0.2), env=env(), lw=True)
    case("lw_fluid_first30", vxa.dense_lattice(2, 23, materials=(3, 4)), 30, 1, sim=S(init_cm_time=0.0005, stop_condition_value=0.2), env=env(), lw=True)
    # the evosoro default for swimmers: fluid AND SelfColEnabled=1 (evosoro/base.py:24); the pincer makes real contacts
    case("lw_fluid_col_pincer", pincer(), 800, 50, sim=S(init_cm_time=0.005, stop_condition_value=0.5, self_col_enabled=1), env=env(), lw=True, col=True)
    case("lw_fluid_dense3_full", vxa.dense_lattice(3, 24, materials=(1, 3, 4)), 0, 1, sim=S(init_cm_time=0.02, stop_condition_value=0.3), env=env(), lw=True, full=True)


def bc_cases():
    """fixed / forced regions (VX_Sim.cpp:517-582, VXS_Voxel.cpp:174-221, :540, :668-672)"""
    S, E, R = vxa.SimParams, vxa.EnvParams, vxa.Region
    free = lambda: E(grav_enabled=0, floor_enabled=0, temp_enabled=0, lattice_dim=0.001)
    soft = [vxa.Material("1MPa", 1e6, density=1.0)]
    # P4 of SURVEY 8c: 20x1x1 cantilever of VX_Benchmark.cpp:36-79, clamped at x = 0, transverse tip load, run to rest
    n = 20
    beam = np.ones((n, 1, 1), dtype=np.int32)
    bcs = [R(0, (0, 0, 0), (0.5 / n, 1, 1), dof_fixed=0x3F), R(0, ((n - 0.5) / n, 0, 0), (0.5 / n, 1, 1), force=(0, 0, -3e-5))]
    case("bc_cantilever20", beam, 80000, 4000, sim=S(bond_damping_z=1.0, slow_damping_z=0.001, init_cm_time=0.0, stop_condition_type=1, stop_condition_value=80000),
         env=free(), palette=soft, bcs=bcs)
    # mixed constraints on a heterogeneous block under gravity: a clamped corner column with a prescribed displacement and
    # rotation, a plane on rollers (z only), a sphere-shaped forced region with force + torque, overlapping regions
    blk = vxa.dense_lattice(4, 41)
    bcs = [R(0, (0, 0, 0), (0.25, 0.25, 1), dof_fixed=0x3F, displace=(1e-4, 0, 2e-4), ang_displace=(0, 0.05, 0.02)),
           R(0, (0.75, 0, 0), (0.25, 1, 0.25), dof_fixed=0x04, displace=(0, 0, -1e-4)),
           R(2, (0.5, 0.5, 1.0), radius=0.3, force=(0.2, -0.1, 0.3), torque=(1e-4, 0, -2e-4)),
           R(0, (0, 0.75, 0.5), (0.5, 0.25, 0.5), dof_fixed=0x38, ang_displace=(0.01, 0, 0)),
           R(1, (0.4, 0.4, 0.0), (0, 0, 1.0), radius=0.2, force=(0, 0, 0.5))]
    case("bc_block4_mixed", blk, 1500, 100, sim=S(init_cm_time=0.01, stop_condition_value=0.2), env=E(floor_enabled=0), bcs=bcs)
    # the first 30 steps of the same, every step
    case("bc_block4_first30", blk, 30, 1, sim=S(init_cm_time=0.0005, stop_condition_value=0.2), env=E(floor_enabled=0), bcs=bcs)


def dev_cases():
    """per-voxel development and evolved stiffness (VX_Sim.cpp:878-991, VXS_Voxel.cpp:224-336)"""
    S, E = vxa.SimParams, vxa.EnvParams
    rng = np.random.RandomState(8)
    u = lambda n, lo, hi: rng.uniform(lo, hi, (n, n, n))
    soft = (1, 3, 4)   # no bone: dt = 6.4e-4 s, so a 0.3 s life is ~470 steps
    # everything at once: initial / final size, growth window, developing phase and amplitude; MinGrowthTime clamps some windows
    pv = {"InitialVoxelSize": u(4, -1, 1), "FinalVoxelSize": u(4, -1, 1), "GrowthTime": u(4, 0, 1), "StartGrowthTime": u(4, 0, 1),
          "FinalPhaseOffset": u(4, -0.5, 0.5), "TempAmpDamp": u(4, 0.2, 1), "FinalTempAmpDamp": u(4, 0.2, 1)}
    case("dev_growth4", vxa.dense_lattice(4, 51, materials=soft), 480, 40, sim=S(init_cm_time=0.02, stop_condition_value=0.3),
         env=E(growth_amplitude=0.75, extra_tags={"MinGrowthTime": 0.05}), phase=u(4, -0.5, 0.5), per_voxel=pv)
    # final size only (growth starts at InitCmTime and lasts the whole life) with a mid-life freeze
    case("dev_freeze3", vxa.dense_lattice(3, 52, materials=soft), 480, 40, sim=S(init_cm_time=0.02, stop_condition_value=0.3, mid_life_freeze_time=0.06),
         env=E(growth_amplitude=0.5), per_voxel={"FinalVoxelSize": u(3, -1, 1)})
    # initial size only: prenatal ramp to a per-voxel size, actuation scaled by size; shrinking below MinTempFact is clamped
    case("dev_initsize3_first40", vxa.dense_lattice(3, 53, materials=soft), 40, 1, sim=S(init_cm_time=0.01, stop_condition_value=0.3, min_temp_fact=0.6),
         env=E(growth_amplitude=0.75), per_voxel={"InitialVoxelSize": u(3, -1, 1)})
    # evolved stiffness: floor contact and contact bonds follow it, the internal bonds keep the palette modulus
    stiff = lambda n: 10 ** rng.uniform(np.log10(5e6), np.log10(5e8), (n, n, n))
    # (_voxcad keeps dt at the palette's bond stiffness, so a much stiffer evolved floor contact chatters and amplifies rounding
    # differences between any two fp64 evaluation orders: the fixture stays within 3x of the palette modulus)
    case("dev_stiffness3", vxa.dense_lattice(3, 54, materials=soft), 600, 50, sim=S(init_cm_time=0.02, stop_condition_value=0.3),
         per_voxel={"Stiffness": 10 ** rng.uniform(np.log10(1e6), np.log10(1.5e7), (3, 3, 3))})
    case("dev_stiffness_col_pincer", pincer(), 600, 50, sim=S(init_cm_time=0.005, stop_condition_value=0.5, self_col_enabled=1), col=True,
         per_voxel={"Stiffness": 10 ** rng.uniform(np.log10(5e6), np.log10(5e7), pincer().shape)})
    # whole CLI run of a developing robot with evolved stiffness: the result file both CLI seams have to reproduce
    rng2 = np.random.RandomState(9)
    u2 = lambda lo, hi: rng2.uniform(lo, hi, (3, 3, 3))
    case("dev_growth3_full", vxa.dense_lattice(3, 56, materials=soft), 0, 1, sim=S(init_cm_time=0.05, stop_condition_value=0.4),
         env=E(growth_amplitude=0.5, extra_tags={"MinGrowthTime": 0.02}), phase=u2(-0.5, 0.5), full=True,
         per_voxel={"InitialVoxelSize": u2(-1, 1), "FinalVoxelSize": u2(-1, 1), "GrowthTime": u2(0, 1), "StartGrowthTime": u2(0, 1),
                    "FinalPhaseOffset": u2(-0.5, 0.5), "Stiffness": 10 ** u2(np.log10(2e6), np.log10(1.2e7))})
    case("lw_stiffness3", vxa.dense_lattice(3, 55, materials=soft), 600, 50, sim=S(init_cm_time=0.02, stop_condition_value=0.3), lw=True, per_voxel={"Stiffness": stiff(3)})


def evosoro_cases():
    """The wire format as evosoro itself writes it: two individuals of the first population of its stock experiment
    (examples/basic.py: 6^3 workspace, CPPN morphologies, SelfColEnabled=1, 5 s lifetime), written by evosoro's own
    write_voxelyze_file (tools/make_evosoro_vxa.py), whole CLI runs."""
    import re
    import make_evosoro_vxa
    files = make_evosoro_vxa.main(os.path.join(TMP, "evosoro_basic"), 6)
    for idx in (1, 2):
        text = open(files[idx]).read()
        fit = os.path.join(TMP, "evosoro_basic_id%d_result.xml" % idx)
        text = re.sub(r"<FitnessFileName>[^<]*</FitnessFileName>", "<FitnessFileName>" + fit + "</FitnessFileName>", text)
        case_text("evosoro_basic_id%d_full" % idx, text, 0, 1, full=True, col=True, fitness_file=fit)


REF = "/root/reference/evosoro"
EXAMPLES = [   # (fixture name, file, land/water fork, short-horizon steps, frame spacing)
    ("ex_example_1", "_voxcad/voxelyzeMain/Example_1.vxa", False, 1000, 250),
    ("ex_example_2", "_voxcad/voxelyzeMain/Example_2.vxa", False, 800, 400),
    ("ex_example_3", "_voxcad/voxelyzeMain/Example_3.vxa", False, 800, 400),
    ("ex_no_actuation", "_voxcad/voxelyzeMain/Example_no_actuation.vxa", False, 1000, 250),
    ("ex_phase_offset", "_voxcad/voxelyzeMain/Example_withPhaseOffset.vxa", False, 1000, 250),
    # <Structure Compression="ZLIB"> + a fixed region: read by the reference built with USE_ZLIB_COMPRESSION (oracle/Makefile: *_zlib)
    ("ex_sticky_floor_bc", "_voxcad/voxelyzeMain/Example_stickyFloorBC.vxa", False, 800, 400),
    ("ex_lw_flapping", "_voxcad_land_water/sample_vxa/flapping_passive_tissue.vxa", True, 1000, 250),
    ("ex_lw_full_cross", "_voxcad_land_water/sample_vxa/full_cross.vxa", True, 800, 400),
    ("ex_lw_hexapus", "_voxcad_land_water/sample_vxa/hexapus.vxa", True, 1000, 250),
    ("ex_lw_hopper", "_voxcad_land_water/sample_vxa/hopper.vxa", True, 1200, 300),
    ("ex_lw_land_stiffness", "_voxcad_land_water/sample_vxa/land_stiffness_phaseoffset.vxa", True, 1000, 250),
    ("ex_lw_pusher", "_voxcad_land_water/sample_vxa/pusher_land.vxa", True, 1200, 300),
    ("ex_lw_quadruped", "_voxcad_land_water/sample_vxa/quadruped_land.vxa", True, 1200, 300),
    ("ex_lw_quadrupus", "_voxcad_land_water/sample_vxa/quadrupus_passive_tissue.vxa", True, 1000, 250),
]


def example_cases(only=None):
    """The inputs the reference ships (voxelyzeMain/Example_*.vxa and _voxcad_land_water/sample_vxa/*.vxa), as shipped
    (SelfColEnabled=1, evolved morphologies, per-voxel phase offsets / stiffness, fluid environments): a short-horizon
    trajectory (<name>.npz) and the whole CLI run with the reference CLI's own result file (<name>_full.npz, final state and
    fitness block only).  Only <FitnessFileName> is redirected."""
    import re
    for name, rel, lw, n_steps, every in EXAMPLES:
        if only and name not in only:
            continue
        text = open(os.path.join(REF, rel), encoding="latin-1").read()
        fit = os.path.join(TMP, name + "_result.xml")
        text = re.sub(r"<FitnessFileName>[^<]*</FitnessFileName>", "<FitnessFileName>" + fit + "</FitnessFileName>", text)
        if "<FitnessFileName>" not in text:   # (Example_stickyFloorBC.vxa names no result file: it gets the <GA> block of the other examples)
            text = text.replace("</Simulator>", "<GA><FitnessFileName>" + fit + "</FitnessFileName><WriteFitnessFile>true</WriteFitnessFile></GA>\n  </Simulator>", 1)
        suffix = "_zlib" if 'Compression="ZLIB"' in text else ""
        case_text(name, text, n_steps, every, lw=lw, col=True, suffix=suffix)
        case_text(name + "_full", text, 0, 1, lw=lw, full=True, col=True, fitness_file=fit, suffix=suffix)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "density":
        pal = vxa.four_material_palette()
        pal[0].density, pal[1].density, pal[2].density, pal[3].density = 1e6, 3e6, 5e5, 2e6
        case("dense3_density", vxa.dense_lattice(3, 14), 1200, 100, sim=vxa.SimParams(init_cm_time=0.02, stop_condition_value=0.2), palette=pal)
    elif len(sys.argv) > 1 and sys.argv[1] == "evosoro":
        evosoro_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "examples":
        example_cases(sys.argv[2:])
    elif len(sys.argv) > 1 and sys.argv[1] == "big":
        case("dense10_s3_full", vxa.dense_lattice(10, 3), 0, 1, full=True)
    elif len(sys.argv) > 1 and sys.argv[1] == "dev":
        dev_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "bc":
        bc_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "fluid":
        fluid_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "col":
        collision_cases()
    else:
        main()
        fluid_cases()
        collision_cases()
        bc_cases()
        dev_cases()
        evosoro_cases()
        example_cases()
